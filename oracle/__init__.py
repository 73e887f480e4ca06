"""TEST INFRASTRUCTURE: CPU oracles for the hot path (see oracle/README.md).  Never imported by the product."""
