"""Import stub (test infrastructure): utils/common_utils.py imports two diffusers names it only uses for the ModelScope path."""
