"""Seeded inputs of tests/golden/make_pipeline_golden.py, importable without /root/reference."""
import numpy as np

SEED_W, SEED_P = 31, 32


def weights(B, H, W):
    return np.random.default_rng(SEED_W).normal(size=(B, H, W)).astype(np.float32)


def target(B, H, W):
    return (np.random.default_rng(SEED_P).random(size=(B, H, W)) * 1.3).astype(np.float32)
