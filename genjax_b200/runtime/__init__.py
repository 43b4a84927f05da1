"""Native runtime plumbing: nvcc driver, C-ABI loader, multi-GPU sharding."""
