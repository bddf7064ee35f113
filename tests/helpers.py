"""Shared test inputs (seeded) for the oracle / CUDA parity tests."""
import numpy as np


def random_surface_cloud(rng, n, extent=12.0, planes=6):
    """Points on a few random planes (keeps ~5-12 neighbours per voxel like LiDAR surfaces)."""
    pts = []
    per = n // planes
    for _ in range(planes):
        origin = rng.uniform(-extent, extent, 3)
        u, v = rng.standard_normal(3), rng.standard_normal(3)
        u /= np.linalg.norm(u)
        v -= u * (u @ v)
        v /= np.linalg.norm(v)
        a, b = rng.uniform(-extent / 2, extent / 2, (2, per))
        pts.append(origin + a[:, None] * u + b[:, None] * v + rng.normal(0, 0.01, (per, 3)))
    return np.concatenate(pts).astype(np.float32)


def random_voxels(rng, n, span=40, batch=2):
    """Unique int32 (b,x,y,z) rows in random order, clustered enough to have neighbours."""
    c = np.concatenate([rng.integers(0, batch, (n, 1)), rng.integers(-span, span, (n, 2)),
                        rng.integers(-4, 4, (n, 1))], axis=1).astype(np.int32)
    _, first = np.unique(c, axis=0, return_index=True)
    c = c[np.sort(first)]
    # batch-sorted like a collated batch
    return c[np.argsort(c[:, 0], kind="stable")]


def record(name, **values):
    """Append the measured numbers of a parity test to gpurun_out/r02_parity_measured.jsonl (the table in
    profiles/ and the bars in the tests come from these lines, not from guesses)."""
    import json
    import os
    root = os.environ.get("GRAFT_REPO_ROOT") or os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    try:
        os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
        with open(os.path.join(root, "gpurun_out", "r02_parity_measured.jsonl"), "a") as f:
            f.write(json.dumps(dict(test=name, **values)) + "\n")
    except OSError:
        pass
