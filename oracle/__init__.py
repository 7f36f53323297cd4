"""CPU oracle -- TEST INFRASTRUCTURE ONLY (see jpd_oracle.c).  Never imported by jacobi_pd_b200."""
