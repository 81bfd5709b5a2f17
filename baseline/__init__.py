"""Stand-in baselines that run on the GPU next to the product (bench.py extras).  Nothing here is shipped or imported
by the package; nothing here is the reference either -- see eager_torch_heads.py."""
