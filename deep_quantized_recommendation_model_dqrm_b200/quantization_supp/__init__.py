"""Drop-in for the reference package ``quantization_supp`` (hot-path subset)."""
