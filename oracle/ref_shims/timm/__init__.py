"""Import shim (test infrastructure only): the reference pins timm 0.9.12, absent from this image."""
