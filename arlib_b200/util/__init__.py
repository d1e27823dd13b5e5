"""Host-side mirror of the reference's ``util`` package (same module names and
signatures) for the graph-CF hot path."""
