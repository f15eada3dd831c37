"""CPU oracle of the reference EM-GMM path -- test infrastructure only (see emgmm_oracle.c header)."""
