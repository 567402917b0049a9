"""oracle/_ref: compiled reference natives (built by oracle/build_ref.py)."""
