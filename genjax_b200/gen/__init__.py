"""Modeling language: capture, code generation and the GFI on fused kernels."""
