"""Test infrastructure (the checker): see oracle/README.md.  Not part of the product."""
