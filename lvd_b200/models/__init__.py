"""Host-side mirror of the reference's `models` package for the hot path: same module names, class names, call
signatures and error behaviour; the arithmetic runs in liblvd_b200.so."""
