"""Test-only stand-in for the absent `tiny_tf` package."""
