"""numpy (fp32, per-operation IEEE) restatements used ONLY as checkers for quantities the reference itself does not
define: the central-difference gradient normals (north_star).  The field values that drive it
come from the reference's own Evaluator through oracle/_ref or from the golden fixtures."""
import numpy as np

f32 = np.float32

EDGE_A = [0, 1, 2, 3, 4, 5, 6, 7, 0, 1, 2, 3]   # marching_lookup.h:10-23
EDGE_B = [1, 2, 3, 0, 5, 6, 7, 4, 4, 5, 6, 7]
CORNER = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]  # marching.cpp:471-472


def apron_coords(coords, step):
    """cs[v+1] = c[v] for v in [-1, M+1]: the reference loop values plus one fp32 step on either side."""
    c = np.asarray(coords, f32)
    step = f32(step)
    return np.concatenate([[f32(c[0] - step)], c, [f32(c[-1] + step)]]).astype(f32)


def soup_gradient_normals(field_ext, cs, iso, cubes, tri_rows):
    """field_ext[z,y,x] on the apron grid (index = vertex+1); cubes = list of (i,j,k,code,tidx) in emission order;
    tri_rows[tidx] = list of edge indices.  Returns normals [T,3,3] fp32."""
    F = np.asarray(field_ext, f32)
    iso = f32(iso)
    out = []
    with np.errstate(all="ignore"):
        for (i, j, k, code, tidx) in cubes:
            val, grad = [], []
            for (dx, dy, dz) in CORNER:
                x, y, z = i + 1 + dx, j + 1 + dy, k + 1 + dz
                val.append(F[z, y, x])
                # the product's definition: the difference times the fp32 reciprocal of the coordinate difference
                gx = f32(f32(F[z, y, x + 1] - F[z, y, x - 1]) * f32(f32(1.0) / f32(cs[x + 1] - cs[x - 1])))
                gy = f32(f32(F[z, y + 1, x] - F[z, y - 1, x]) * f32(f32(1.0) / f32(cs[y + 1] - cs[y - 1])))
                gz = f32(f32(F[z + 1, y, x] - F[z - 1, y, x]) * f32(f32(1.0) / f32(cs[z + 1] - cs[z - 1])))
                grad.append((gx, gy, gz))
            enrm = {}
            for e in range(12):
                a, b = EDGE_A[e], EDGE_B[e]
                if ((code >> a) ^ (code >> b)) & 1 == 0:
                    continue
                if sum(CORNER[a]) > sum(CORNER[b]):
                    a, b = b, a   # defined on the grid edge: from its lower end point to the upper one, for every cube sharing it
                f1, f2 = val[a], val[b]
                t = f32(f32(iso - f1) / f32(f2 - f1))
                if np.isinf(t) or np.isnan(t):
                    t = f32(0.5)
                n = [f32(grad[a][q] + f32(t * f32(grad[b][q] - grad[a][q]))) for q in range(3)]
                s = f32(f32(f32(n[0] * n[0]) + f32(n[1] * n[1])) + f32(n[2] * n[2]))
                inv = f32(f32(1.0) / np.sqrt(s, dtype=f32))
                enrm[e] = [f32(n[q] * inv) for q in range(3)]
            row = tri_rows[tidx]
            for t0 in range(0, len(row), 3):
                out.append([enrm[row[t0]], enrm[row[t0 + 1]], enrm[row[t0 + 2]]])
    return np.array(out, f32).reshape(-1, 3, 3)
