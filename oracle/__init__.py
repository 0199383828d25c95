"""CPU oracle -- TEST INFRASTRUCTURE ONLY (see oracle/ccv2_oracle.h). Never imported by the product package."""
