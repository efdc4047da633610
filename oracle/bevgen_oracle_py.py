"""Second, independent restatement of the reference hot path in plain Python loops over numpy float32 scalars.

TEST INFRASTRUCTURE ONLY.  It exists to pin oracle/bevgen_oracle.c: the two were written separately from the
reference source (each line cites it) and tests/test_oracle_cross.py requires them to agree bit for bit on
seeded inputs.  It is slow (pure Python), so it is only run on small sensor shapes.

Reference: soytony/Point-Cloud-Preprocessing-Tools @ d94040e, BatchMultiBevGen.cpp / BatchMultiBevGen.h.
"""
import ctypes
import ctypes.util
import math

import numpy as np

f32 = np.float32
_libm = ctypes.CDLL(ctypes.util.find_library("m"))
_libm.atan2f.restype = ctypes.c_float
_libm.atan2f.argtypes = [ctypes.c_float, ctypes.c_float]

INT_MIN = -2 ** 31


def cvtt(v):
    """x86 cvttsd2si / cvttss2si: truncate toward zero, INT_MIN when NaN or out of int32 range."""
    v = float(v)
    if not (-2147483649.0 < v < 2147483648.0):
        return INT_MIN
    return int(v)  # Python int() truncates toward zero


def c_round(v):
    """C round(): half away from zero, on a double."""
    v = float(v)
    if math.isnan(v) or math.isinf(v):
        return v
    return math.copysign(math.floor(abs(v) + 0.5), v) if abs(v) < 2 ** 52 else v


def ordered_cloud(N, H, x, y, z, inten, row, col, label):
    """getOrderedCloud, BatchMultiBevGen.cpp:94-117."""
    S = N * H
    o = dict(x=np.zeros(S, f32), y=np.zeros(S, f32), z=np.zeros(S, f32), intensity=np.zeros(S, f32),
             label=np.zeros(S, np.int16), owner=np.zeros(S, np.uint32))   # :98 resize() value-initialises
    for i in range(len(x)):                                                 # :102 serial => last writer wins
        r, c = int(row[i]), int(col[i])
        if r < 0 or r >= N:                                                 # :106
            continue
        if c < 0 or c >= H:                                                 # :109
            continue
        p = r * H + c                                                       # :113
        o["x"][p], o["y"][p], o["z"][p], o["intensity"][p], o["label"][p] = x[i], y[i], z[i], inten[i], label[i]
        o["owner"][p] = i + 1
    return o


def belonging_grid(px, py):
    """getBelongingGrid, BatchMultiBevGen.h:73-99."""
    nx = f32(float(px) + 75.0)          # :78 float + double literal, stored to float
    ny = f32(float(py) + 50.0)          # :79
    sr = cvtt(math.floor(float(nx) / 2.0)) if math.isfinite(float(nx)) else INT_MIN   # :81
    sc = cvtt(math.floor(float(ny) / 2.0)) if math.isfinite(float(ny)) else INT_MIN   # :82
    sr = 74 if sr >= 75 else sr
    sr = 0 if sr < 0 else sr
    sc = 49 if sc >= 50 else sc
    sc = 0 if sc < 0 else sc
    return sr, sc


def c_rem(a, b):
    """C++ '%' on ints (truncated division: result takes the sign of the dividend)."""
    return int(math.fmod(a, b))


def mark_ground(N, H, G, o):
    """markGroundPoints, BatchMultiBevGen.cpp:119-252.  Returns (label, gm_after_loop1, gm_final, avg)."""
    X, Y, Z, I = o["x"], o["y"], o["z"], o["intensity"]
    label = o["label"].copy()
    gm = np.zeros((N, H), np.int8)                                          # :123
    avg = np.zeros((75, 50), f32)                                           # :133
    num = np.full((75, 50), f32(0.01), f32)                                 # :135
    for c in range(H):                                                      # :139
        for r in range(N - 1, N - G - 1, -1):                               # :140
            lower = r * H + c
            upper = (r - 1) * H + c
            if I[upper] == -1:                                              # :146
                upper = (r - 1) * H + c_rem(c + 2, H)
            if I[upper] == -1:                                              # :151
                upper = (r - 1) * H + c_rem(c - 2, H)
            if I[upper] == -1 and r >= 2:                                   # :157
                upper = (r - 2) * H + c
            if I[lower] == -1 or I[upper] == -1:                            # :162
                gm[r, c] = -1
                continue
            dx = f32(X[upper] - X[lower]); dy = f32(Y[upper] - Y[lower]); dz = f32(Z[upper] - Z[lower])   # :169-171
            hyp = np.sqrt(f32(f32(dx * dx) + f32(dy * dy)))                 # sqrtf, correctly rounded
            angle = f32(float(_libm.atan2f(float(dz), float(hyp))) * 180.0 / math.pi)   # :173
            if abs(f32(angle - f32(0.0))) <= f32(10.0):                     # :179
                gm[r, c] = 1
                gm[r - 1, c] = 1
    gm1 = gm.copy()
    for r in range(N):                                                      # :187
        for c in range(H):
            if gm[r, c] != 1:
                continue
            p = r * H + c
            sr, sc = belonging_grid(X[p], Y[p])
            avg[sr, sc] = f32(avg[sr, sc] + Z[p])                           # :198
            num[sr, sc] = f32(num[sr, sc] + f32(1))                         # :205
    with np.errstate(all="ignore"):
        avg = (avg / num).astype(f32)                                       # :210
    nbs = ((-1, 0), (0, 1), (0, -1), (1, 0))                                # :73-84
    for r in range(N):                                                      # :216
        for c in range(H):
            p = r * H + c
            sr, sc = belonging_grid(X[p], Y[p])                             # :223
            for dr, dc in nbs:
                nr, nc = sr + dr, sc + dc
                if nr < 0 or nr >= 75 or nc < 0 or nc >= 50:                # :231
                    continue
                with np.errstate(all="ignore"):
                    d = f32(Z[p] - avg[nr, nc])
                if float(d) > 0.30:                                         # :236-237
                    gm[r, c] = 0
                    break
            if gm[r, c] == 1:                                               # :244
                label[p] = 0
    return label, gm1, gm, avg


def _cell(v):
    w = f32(f32(v) + f32(112)) / f32(1.0)                                   # :279 (pi.x + MAX_RANGE) / interval
    return cvtt(c_round(float(f32(w)) + 0.5))


def bevs(N, H, height_res, o, label):
    """Binning of computeAndSaveMultiBev :278-292 and computeAndSaveSingleBev :342-356."""
    multi = np.zeros((24, 224, 224), np.uint8)
    single = np.zeros((224, 224), np.uint8)
    hr = f32(height_res)
    for p in range(N * H):
        with np.errstate(all="ignore"):
            x = _cell(o["x"][p]); y = _cell(o["y"][p])
            layer = cvtt(c_round(float(f32(f32(o["z"][p] / hr) + f32(2.0)))))   # :281
            h = cvtt(float(f32(o["z"][p] + f32(2.0))) * 4.0)                    # :345
        h = min(max(0, h), 255)                                                 # :346
        inside = not (x < 0 or x >= 224 or y < 0 or y >= 224)
        if inside and label[p] != 0:
            if single[x, y] < h:                                                # :353
                single[x, y] = h
            if 0 <= layer < 24:                                                 # :284
                multi[layer, x, y] = 255
    return single, multi


def d2(q, m):
    """nanoflann L2_Adaptor::evalMetric for dim 3 (nanoflann.hpp:383-407)."""
    r = f32(0)
    for k in range(3):
        d = f32(q[k] - m[k])
        r = f32(r + f32(d * d))
    return r


def select_major(xyz):
    """selectMajorFrames, BatchMultiBevGen.cpp:502-566 (exhaustive 1-NN instead of the KD-tree)."""
    xyz = np.asarray(xyz, f32).reshape(-1, 3)
    majors = [0]
    for i in range(1, len(xyz)):
        last = xyz[majors[-1]]
        dd = [f32(xyz[i][k] - last[k]) for k in range(3)]
        dist = np.sqrt(f32(f32(f32(dd[0] * dd[0]) + f32(dd[1] * dd[1])) + f32(dd[2] * dd[2])))   # Utility.cpp:43-49
        if dist < f32(20.0):
            continue
        best = min(d2(xyz[i], xyz[m]) for m in majors)
        if best < f32(400.0):
            continue
        majors.append(i)
    return np.array(majors, np.int32)


def labels(xyz, majors):
    """getKeyFrameLabel, BatchMultiBevGen.cpp:575-636 (exhaustive 2-NN, first-found wins ties)."""
    xyz = np.asarray(xyz, f32).reshape(-1, 3)
    K, M = len(xyz), len(majors)
    out = np.zeros((K, M), f32)
    for i in range(K):
        cand = [0, 0]; dist = [f32(0), f32(np.finfo(f32).max)]
        count = 0
        for j, m in enumerate(majors):                      # KNNResultSet::addPoint, nanoflann.hpp:175-203
            d = d2(xyz[i], xyz[m])
            k = count
            while k > 0 and dist[k - 1] > d:
                if k < 2:
                    dist[k], cand[k] = dist[k - 1], cand[k - 1]
                k -= 1
            if k < 2:
                dist[k], cand[k] = d, j
            if count < 2:
                count += 1
        if i == majors[cand[0]]:
            out[i, cand[0]] = 1.0
        else:
            w0 = f32(1.0 / (float(dist[0]) + 1e-5)); w1 = f32(1.0 / (float(dist[1]) + 1e-5))
            s = f32(w0 + w1)
            out[i, cand[0]] = f32(w0 / s)
            out[i, cand[1]] = f32(w1 / s)
    return out
