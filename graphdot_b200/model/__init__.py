"""Callers of the marginalized-kernel hot path that continue on the device
(SURVEY section 8f)."""
