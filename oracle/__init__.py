"""ORACLE — test infrastructure only (see oracle/ecfft_oracle.h).  Never imported by ecfft_b200/."""
