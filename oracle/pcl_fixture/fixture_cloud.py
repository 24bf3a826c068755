"""The seeded cloud of the PCL pinning kit, generated with integer / IEEE +,-,*,/ arithmetic only, so that the C++ program
(dump_pcl_fixture.cpp, same recurrence) and numpy produce the same floats bit for bit on any platform.

Three surfaces (floor, wall, 30-degree ramp) with triangular noise, 20 000 points, insertion order = generation order
interleaved over the surfaces (drives PCL's dynamic bounding box through several growth epochs)."""
import numpy as np

N = 20_000
MASK = (1 << 32) - 1


def lcg_stream(n, seed=12345):
    """n uniform doubles in [0, 1): x <- 1664525 x + 1013904223 mod 2^32, u = x / 2^32"""
    out = np.empty(n, np.float64)
    x = seed
    for i in range(n):
        x = (1664525 * x + 1013904223) & MASK
        out[i] = x / 4294967296.0
    return out


def cloud():
    u = lcg_stream(N * 5).reshape(N, 5)
    pts = np.empty((N, 3), np.float64)
    for i in range(N):
        a, b, n1, n2, n3 = u[i]
        noise = 0.004 * np.array([n1 + n2 - 1.0, n2 + n3 - 1.0, n3 + n1 - 1.0])
        s = i % 3
        if s == 0:
            p = np.array([0.3 + 4.0 * a, 0.2 + 3.0 * b, 0.1])
        elif s == 1:
            p = np.array([0.3 + 4.0 * a, 3.2, 0.1 + 2.5 * b])
        else:
            p = np.array([4.3 + 2.0 * a, 0.2 + 3.0 * b, 0.1 + 1.1547 * a])
        pts[i] = p + noise
    return pts.astype(np.float32)


if __name__ == "__main__":
    import sys
    cloud().tofile(sys.argv[1] if len(sys.argv) > 1 else "fixture_cloud.f32")
