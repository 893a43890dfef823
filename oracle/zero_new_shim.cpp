// TEST INFRASTRUCTURE ONLY — linked into oracle/_ref/extract_ref_z (the parity build of the
// unmodified reference) and nowhere else.
//
// The reference reads heap memory it never wrote (SURVEY.md Appendix A, Q5/Q6:
// record_ref_hit / record_ref_index tails, src/extract_ref_normal_peak.cpp:931-945 and :247,262).
// Replacing the array forms of operator new with zero-filling versions makes those reads
// deterministic (what a fresh mmap gives) without touching any reference source line.
#include <cstdlib>
#include <new>

void* operator new[](std::size_t n) {
    void* p = std::calloc(n ? n : 1, 1);
    if (!p) throw std::bad_alloc();
    return p;
}
void operator delete[](void* p) noexcept { std::free(p); }
void operator delete[](void* p, std::size_t) noexcept { std::free(p); }
