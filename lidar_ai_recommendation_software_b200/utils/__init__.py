"""Drop-in for the reference's `utils` package (hot-path module only: data_processing)."""
