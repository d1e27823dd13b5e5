"""Drop-in graph-CF recommenders (same module / class names as the reference's
``recommender`` package): LightGCN, NGCF, SimGCL, XSimGCL."""
