"""localhgt_b200 — B200-native k-mer screening stage of LocalHGT (`extract_ref`)."""
