"""Import stub for rawpy (absent here; RAW photo loading is out of scope)."""
