"""Minimal PCD v0.7 (binary) reader/writer for pcl::PointXYZIRCT clouds — test/tooling helper (numpy only).
Layout as written by pcl::io::savePCDFileBinary for the point type of BatchMultiBevGen.h:43-66."""
import numpy as np

DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("intensity", "<f4"), ("row", "<u2"), ("col", "<u2"),
                  ("t", "<u4"), ("label", "<i2")])
assert DTYPE.itemsize == 26


def header(n, width=None, height=1):
    width = n if width is None else width
    return ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity row col t label\n"
            "SIZE 4 4 4 4 2 2 4 2\nTYPE F F F F U U U I\nCOUNT 1 1 1 1 1 1 1 1\nWIDTH %d\nHEIGHT %d\n"
            "VIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA binary\n" % (width, height, n)).encode()


def records(f):
    n = len(f["x"])
    rec = np.zeros(n, DTYPE)
    for k in DTYPE.names:
        if k in f:
            rec[k] = f[k]
    return rec


def write(path, f, **kw):
    rec = records(f)
    with open(path, "wb") as fp:
        fp.write(header(len(rec), **kw))
        fp.write(rec.tobytes())


def read(path):
    raw = open(path, "rb").read()
    i = raw.index(b"DATA binary\n") + len(b"DATA binary\n")
    hdr = raw[:i].decode()
    n = int([l for l in hdr.splitlines() if l.startswith("POINTS")][0].split()[1])
    rec = np.frombuffer(raw[i:i + n * 26], DTYPE)
    return {k: rec[k].copy() for k in DTYPE.names}, hdr


def write_ascii(path, f, fields=("x", "y", "z", "intensity", "row", "col", "t", "label")):
    n = len(f["x"])
    size = {"x": 4, "y": 4, "z": 4, "intensity": 4, "row": 2, "col": 2, "t": 4, "label": 2}
    typ = {"x": "F", "y": "F", "z": "F", "intensity": "F", "row": "U", "col": "U", "t": "U", "label": "I"}
    with open(path, "w") as fp:
        fp.write("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS %s\nSIZE %s\nTYPE %s\nCOUNT %s\nWIDTH %d\nHEIGHT 1\n"
                 "VIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA ascii\n" % (" ".join(fields), " ".join(str(size[k]) for k in fields),
                                                                      " ".join(typ[k] for k in fields), " ".join("1" for _ in fields), n, n))
        for i in range(n):
            fp.write(" ".join(repr(float(f[k][i])) if typ[k] == "F" else str(int(f[k][i])) for k in fields) + "\n")


def write_binary_layout(path, f, fields=("label", "x", "_", "y", "z", "col", "row", "intensity", "t")):
    """Binary PCD with an arbitrary field order; "_" is a 3-byte padding field (SIZE 1 TYPE U COUNT 3) as PCL writes
    for alignment holes.  Exercises the by-name field mapping and the interleaved-record layout with odd offsets."""
    n = len(f["x"])
    base = {k: DTYPE.fields[k][0] for k in DTYPE.names}
    dt = np.dtype([((k if k != "_" else "pad%d" % i), (base[k] if k != "_" else "V3")) for i, k in enumerate(fields)])
    rec = np.zeros(n, dt)
    for i, k in enumerate(fields):
        if k != "_":
            rec[k] = f[k]
    size = [str(base[k].itemsize) if k != "_" else "1" for k in fields]
    typ = [{"f": "F", "u": "U", "i": "I"}[base[k].kind] if k != "_" else "U" for k in fields]
    cnt = ["1" if k != "_" else "3" for k in fields]
    with open(path, "wb") as fp:
        fp.write(("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS %s\nSIZE %s\nTYPE %s\nCOUNT %s\nWIDTH %d\nHEIGHT 1\n"
                  "VIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA binary\n" % (" ".join(fields), " ".join(size), " ".join(typ), " ".join(cnt), n, n)).encode())
        fp.write(rec.tobytes())
