"""imported for its side effects by losses.py:9; nothing to patch here."""
