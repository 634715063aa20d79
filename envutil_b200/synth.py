"""Deterministic synthetic scenes.

The value of a texel is an analytic function of its viewing DIRECTION d, so that every
representation of the same scene (lat/lon, cubemap, biatan6, rectilinear facet) is
consistent, plus per-texel hash noise that makes interpolation errors visible:

    f_c(d) = 0.5 + 0.25*sin(k1_c*lon + phi_c)*cos(lat) + 0.2*sin(k2_c*lat) + 0.05*n_c(texel)

with (k1,k2,phi) = (3,5,0),(4,7,1),(2,9,2) for R,G,B and n_c uniform in [-1,1] from
splitmix64(seed 0x5eed0000+c, texel index); clamped to [0,1]. Pixel centres follow
envutil's edge-to-edge convention (reference README.md:955-961, stepper.h:324-333).
Coordinate system: x right, y down, z forward (reference envutil_basic.h:66-75).
"""
import numpy as np

_K = ((3.0, 5.0, 0.0), (4.0, 7.0, 1.0), (2.0, 9.0, 2.0))


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def _noise(idx, c):
    with np.errstate(over="ignore"):
        h = _splitmix64(idx.astype(np.uint64) + np.uint64(0x5EED0000 + c) * np.uint64(0x100000001B3))
    return (h >> np.uint64(40)).astype(np.float64) * (2.0 / float(1 << 24)) - 1.0


def scene(lon, lat, texel_index, noise=0.05):
    """Evaluate the scene for directions given as lon/lat (float64 arrays) -> float32 [...,3]."""
    out = np.empty(lon.shape + (3,), dtype=np.float32)
    cl = np.cos(lat)
    for c, (k1, k2, ph) in enumerate(_K):
        v = 0.5 + 0.25 * np.sin(k1 * lon + ph) * cl + 0.2 * np.sin(k2 * lat)
        if noise:
            v = v + noise * _noise(texel_index, c)
        out[..., c] = np.clip(v, 0.0, 1.0)
    return out


def _centres(n, a0, a1):
    i = np.arange(n, dtype=np.float64)
    return a0 + (a1 - a0) * (2.0 * i + 1.0) / (2.0 * n)


def _rows(h, rows_per_chunk=512):
    for y0 in range(0, h, rows_per_chunk):
        yield y0, min(h, y0 + rows_per_chunk)


WORKERS = 1  # row chunks are independent: > 1 evaluates them on that many threads (numpy's ufuncs release the GIL)


def _each(chunks, fn):
    chunks = list(chunks)
    if WORKERS <= 1 or len(chunks) < 2:
        for c in chunks:
            fn(*c)
        return
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=WORKERS) as ex:
        list(ex.map(lambda c: fn(*c), chunks))


def latlon(width, height=None, noise=0.05):
    """Full spherical (360x180) lat/lon image, width x width/2."""
    height = height or width // 2
    out = np.empty((height, width, 3), dtype=np.float32)
    lon = _centres(width, -np.pi, np.pi)
    lat = _centres(height, -np.pi / 2, np.pi / 2)
    def chunk(y0, y1):
        LON, LAT = np.meshgrid(lon, lat[y0:y1])
        idx = (np.arange(y0, y1, dtype=np.int64)[:, None] * width + np.arange(width, dtype=np.int64)[None, :])
        out[y0:y1] = scene(LON, LAT, idx, noise)
    _each(_rows(height, 128 if WORKERS > 1 else 512), chunk)
    return out


def _dir_to_lonlat(x, y, z):
    lon = np.arctan2(x, z)
    lat = np.arctan2(y, np.sqrt(x * x + z * z))
    return lon, lat


def cubemap(face_px, biatan6=False, hfov_deg=90.0, noise=0.05):
    """1:6 vertical cubemap strip (left,right,top,bottom,front,back: reference
    envutil_basic.h:56-64), face orientation as produced by the reference's cubemap
    target (stepper.h:1304-1331). biatan6=True applies the in-plane tan(p*pi/4) transform
    (stepper.h:1492-1493)."""
    w = face_px
    out = np.empty((6 * w, w, 3), dtype=np.float32)
    ext = np.tan(np.radians(hfov_deg) / 2.0)
    p = _centres(w, -ext, ext)
    if biatan6:
        p = np.tan(p * (np.pi / 4.0))
    one = np.ones((1, 1))
    def chunk(face, y0, y1):
            P0, P1 = np.meshgrid(p, p[y0:y1])
            if face == 0:    # left
                x, y, z = -one, P1, P0
            elif face == 1:  # right
                x, y, z = one, P1, -P0
            elif face == 2:  # top
                x, y, z = -P0, -one, -P1
            elif face == 3:  # bottom
                x, y, z = -P0, one, P1
            elif face == 4:  # front
                x, y, z = P0, P1, one
            else:            # back
                x, y, z = -P0, P1, -one
            x, y, z = np.broadcast_arrays(x, y, z)
            lon, lat = _dir_to_lonlat(x, y, z)
            idx = ((face * w + np.arange(y0, y1, dtype=np.int64))[:, None] * w
                   + np.arange(w, dtype=np.int64)[None, :])
            out[face * w + y0: face * w + y1] = scene(lon, lat, idx, noise)
    _each(((face, y0, y1) for face in range(6) for y0, y1 in _rows(w, 128 if WORKERS > 1 else 512)), chunk)
    return out


def rotation(yaw_deg=0.0, pitch_deg=0.0, roll_deg=0.0):
    """3x3 matrix R (float64) with rows = images of e_x,e_y,e_z: camera-frame ray r (row
    vector) -> world ray r @ R. Conventions as the reference's rotate_3d
    (envutil_payload.cc:136-218): roll about z (forward), pitch about x (right), yaw about
    y (down)."""
    r, p, y = np.radians([roll_deg, pitch_deg, yaw_deg])
    ci, cj, ch = np.cos(r / 2), np.cos(p / 2), np.cos(y / 2)
    si, sj, sh = np.sin(r / 2), np.sin(p / 2), np.sin(y / 2)
    cc, cs, sc, ss = ci * ch, ci * sh, si * ch, si * sh
    q = np.empty(4)
    v = np.empty(3)
    v[2] = cj * sc - sj * cs
    v[0] = cj * ss + sj * cc
    v[1] = cj * cs - sj * sc
    qr = cj * cc + sj * ss
    R = np.empty((3, 3))
    for k in range(3):
        e = np.zeros(3)
        e[k] = 1.0
        a = np.cross(v, e)
        b = np.cross(v, a)
        R[k] = e + 2.0 * (qr * a + b)
    return R


def rectilinear_facet(width, height, hfov_deg, yaw_deg=0.0, pitch_deg=0.0, roll_deg=0.0,
                      gain=1.0, noise=0.05, seed_offset=0, rows=None, cols=None):
    """Rectilinear photo of the scene taken by a camera with the given orientation;
    gain scales the linear values before clamping to [0,1] (exposure bracket).
    rows=(r0, r1) / cols=(c0, c1): only that window of the width x height image - the very texels the whole image
    has there (directions and noise are functions of the texel's position in the whole image)."""
    r0, r1 = rows if rows is not None else (0, height)
    c0, c1 = cols if cols is not None else (0, width)
    out = np.empty((r1 - r0, c1 - c0, 3), dtype=np.float32)
    ex = np.tan(np.radians(hfov_deg) / 2.0)
    ey = ex * height / width
    px = _centres(width, -ex, ex)[c0:c1]
    py = _centres(height, -ey, ey)
    R = rotation(yaw_deg, pitch_deg, roll_deg)
    def chunk(y0, y1):
        X, Y = np.meshgrid(px, py[y0:y1])
        Z = np.ones_like(X)
        wx = X * R[0, 0] + Y * R[1, 0] + Z * R[2, 0]
        wy = X * R[0, 1] + Y * R[1, 1] + Z * R[2, 1]
        wz = X * R[0, 2] + Y * R[1, 2] + Z * R[2, 2]
        lon, lat = _dir_to_lonlat(wx, wy, wz)
        idx = (np.arange(y0, y1, dtype=np.int64)[:, None] * width
               + np.arange(c0, c1, dtype=np.int64)[None, :] + seed_offset)
        v = scene(lon, lat, idx, noise).astype(np.float64) * gain
        out[y0 - r0:y1 - r0] = np.clip(v, 0.0, 1.0)
    step = 128 if WORKERS > 1 else 512
    _each(((y0, min(r1, y0 + step)) for y0 in range(r0, r1, step)), chunk)
    return out
