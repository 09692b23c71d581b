"""Host-side mirror of the reference's ``util`` package for the shift-layer path."""
