"""Drop-in for the reference's `models` package: crowd_density_model, crowd_flow_model."""
