"""Shared helpers for the parity tests."""
import json

import numpy as np


def load_meta(golden):
    return json.loads(bytes(golden["meta_json"]).decode())


def configure(ctx, case):
    """Apply a golden case's parameters to an mcb Context through the C ABI."""
    assert ctx.set_equation(case["eq"]) == 0
    assert ctx.set_grid_step(case["step"]) == case["M"]
    ctx.set_scaling(*case["scale"])
    ctx.set_surface_constant(case["iso"])
    for i in range(3):
        ctx.set_constraint(i, ">", 0.0, False)
    for i, (lhs, op, rhs) in enumerate(case.get("cons", [])):
        assert ctx.set_equation(lhs, slot=i + 1) == 0
        assert ctx.set_constraint(i, op, rhs, True) == 0


def same_bits(a, b):
    """fp32 arrays equal bit for bit, NaNs compared as a class."""
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    if a.shape != b.shape:
        return False
    eq = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
    return bool(eq.all())


def rel_close(a, b, tol=1e-5):
    """north_star tolerance: |a-b| <= tol * max(1, |b|) (a purely relative test is ill-posed at 0)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return bool(np.all(np.abs(a - b) <= tol * np.maximum(1.0, np.abs(b))))


