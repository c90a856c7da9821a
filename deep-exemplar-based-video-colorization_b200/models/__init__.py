"""Drop-in replacements for the reference's models.NonlocalNet / models.ColorVidNet (same import paths)."""
