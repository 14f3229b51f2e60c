"""Fused CUDA (sm_100a) anti-aliased activation; see activation1d.py."""
