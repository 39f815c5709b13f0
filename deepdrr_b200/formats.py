"""On-disk CT formats: NIfTI-1 (.nii / .nii.gz) and NRRD (.nrrd / .nhdr) without nibabel / pynrrd.

The reference loads these through third-party readers (``nib.load`` in vol/volume.py:633-641 and ``nrrd.read`` in
vol/volume.py:868-869; both packages are absent here), so this module restates the two published file formats --
NIfTI-1 (nifti1.h, 348-byte header) and NRRD 0004/0005 (teem.sf.net/nrrd/format.html) -- far enough for CT volumes:
3-D (or 4-D with a unit 4th axis) scalar arrays, raw or gzip payload, either byte order.  What the reference takes
from the readers is kept: NIfTI ``affine`` chosen sform -> qform -> pixdim fallback like nibabel, voxels scaled by
``scl_slope / scl_inter`` like ``get_fdata()``; NRRD ``(data, header)`` with ``space directions`` / ``space origin`` as
float arrays like pynrrd.  SURVEY.md 8(f) row 4.
"""
from __future__ import annotations

import gzip
import os
import re
import struct
import zlib
from typing import Dict, Tuple

import numpy as np

# ----------------------------------------------------------------------------------------------- NIfTI-1
_NIFTI_DTYPES = {2: "u1", 4: "i2", 8: "i4", 16: "f4", 64: "f8", 256: "i1", 512: "u2", 768: "u4", 1024: "i8", 1280: "u8"}
_NIFTI_CODES = {np.dtype(v).str[1:]: k for k, v in _NIFTI_DTYPES.items()}


def _open_maybe_gz(path) -> bytes:
    raw = open(path, "rb").read()
    return gzip.decompress(raw) if raw[:2] == b"\x1f\x8b" else raw


def _quaternion_affine(b, c, d, qoff, pixdim) -> np.ndarray:
    """nifti1.h "METHOD 2": rotation from the (b, c, d) quaternion, columns scaled by pixdim[1..3], qfac on k."""
    a2 = 1.0 - (b * b + c * c + d * d)
    if a2 < 1e-7:  # nifti1_io: renormalise, a = 0 (180 degree rotation)
        s = 1.0 / np.sqrt(b * b + c * c + d * d)
        b, c, d, a = b * s, c * s, d * s, 0.0
    else:
        a = np.sqrt(a2)
    r = np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                  [2 * (b * c + a * d), a * a + c * c - b * b - d * d, 2 * (c * d - a * b)],
                  [2 * (b * d - a * c), 2 * (c * d + a * b), a * a + d * d - b * b - c * c]], dtype=np.float64)
    qfac = -1.0 if pixdim[0] < 0 else 1.0
    zooms = np.array([pixdim[1], pixdim[2], pixdim[3] * qfac], dtype=np.float64)
    m = np.eye(4)
    m[:3, :3] = r * zooms[None, :]
    m[:3, 3] = qoff
    return m


def read_nifti(path) -> Tuple[np.ndarray, np.ndarray, Dict]:
    """Returns (voxels float64 [ni, nj, nk] scaled like nibabel's ``get_fdata``, affine 4x4 float64, header dict)."""
    raw = _open_maybe_gz(path)
    if len(raw) < 348:
        raise ValueError(f"{path}: too short for a NIfTI-1 header")
    for e in "<>":
        if struct.unpack(e + "i", raw[0:4])[0] == 348:
            break
    else:
        raise ValueError(f"{path}: not a NIfTI-1 file (sizeof_hdr != 348)")
    magic = raw[344:348]
    if magic not in (b"n+1\0", b"ni1\0"):
        raise ValueError(f"{path}: bad NIfTI-1 magic {magic!r}")
    if magic == b"ni1\0":
        raise ValueError(f"{path}: detached .hdr/.img pairs are not supported; convert to .nii")
    dim = struct.unpack(e + "8h", raw[40:56])
    datatype, bitpix = struct.unpack(e + "2h", raw[70:74])
    pixdim = struct.unpack(e + "8f", raw[76:108])
    vox_offset, slope, inter = struct.unpack(e + "3f", raw[108:120])
    xyzt_units = raw[123]
    qform_code, sform_code = struct.unpack(e + "2h", raw[252:256])
    qb, qc, qd, qx, qy, qz = struct.unpack(e + "6f", raw[256:280])
    srow = np.array(struct.unpack(e + "12f", raw[280:328]), dtype=np.float64).reshape(3, 4)
    if datatype not in _NIFTI_DTYPES:
        raise ValueError(f"{path}: unsupported NIfTI datatype code {datatype}")
    nd = dim[0]
    if not 1 <= nd <= 7:
        raise ValueError(f"{path}: bad dim[0] = {nd}")
    shape = tuple(int(x) for x in dim[1:1 + nd])
    while len(shape) > 3 and shape[-1] == 1:
        shape = shape[:-1]
    if len(shape) != 3:
        raise ValueError(f"{path}: expected a 3-D volume, got shape {shape}")
    dt = np.dtype(e + _NIFTI_DTYPES[datatype])
    n = int(np.prod(shape))
    off = int(vox_offset) if vox_offset >= 352 else 352
    if len(raw) < off + n * dt.itemsize:
        raise ValueError(f"{path}: voxel payload truncated")
    vox = np.frombuffer(raw, dtype=dt, count=n, offset=off).reshape(shape, order="F")
    data = vox.astype(np.float64)
    if np.isfinite(slope) and slope != 0 and not (slope == 1 and inter == 0):   # nibabel: (0 | nan) slope = "no scaling"
        data = data * float(slope) + (float(inter) if np.isfinite(inter) else 0.0)
    if sform_code > 0:
        affine = np.vstack([srow, [0, 0, 0, 1]])
    elif qform_code > 0:
        affine = _quaternion_affine(qb, qc, qd, (qx, qy, qz), pixdim)
    else:  # nibabel get_base_affine(): pixdim scaling, centre of the array at the origin, first axis flipped
        zooms = np.array([abs(pixdim[1]) or 1.0, abs(pixdim[2]) or 1.0, abs(pixdim[3]) or 1.0])
        zooms[0] *= -1
        affine = np.eye(4)
        affine[:3, :3] = np.diag(zooms)
        affine[:3, 3] = -(np.array(shape) - 1) / 2.0 * zooms
    units = {0: "unknown", 1: "meter", 2: "mm", 3: "micron"}.get(xyzt_units & 7, "unknown")
    hdr = {"dim": dim, "datatype": datatype, "bitpix": bitpix, "pixdim": pixdim, "scl_slope": slope, "scl_inter": inter,
           "qform_code": qform_code, "sform_code": sform_code, "xyz_units": units, "endian": e, "dtype": dt}
    return data, affine, hdr


def write_nifti(path, data: np.ndarray, affine: np.ndarray) -> None:
    """Single-file NIfTI-1 (gzip when the name ends in .gz) with the affine in the sform and a matching qform code 0."""
    data = np.asarray(data)
    if data.ndim != 3:
        raise ValueError("write_nifti: 3-D arrays only")
    key = data.dtype.str[1:]
    if key not in _NIFTI_CODES:
        raise ValueError(f"write_nifti: unsupported dtype {data.dtype}")
    affine = np.asarray(affine, dtype=np.float64).reshape(4, 4)
    h = bytearray(352)
    struct.pack_into("<i", h, 0, 348)
    struct.pack_into("<8h", h, 40, 3, data.shape[0], data.shape[1], data.shape[2], 1, 1, 1, 1)
    struct.pack_into("<2h", h, 70, _NIFTI_CODES[key], data.dtype.itemsize * 8)
    zooms = np.sqrt((affine[:3, :3] ** 2).sum(axis=0))
    struct.pack_into("<8f", h, 76, 1.0, zooms[0], zooms[1], zooms[2], 0.0, 0.0, 0.0, 0.0)
    struct.pack_into("<3f", h, 108, 352.0, 1.0, 0.0)
    h[123] = 2  # mm
    struct.pack_into("<2h", h, 252, 0, 2)
    struct.pack_into("<12f", h, 280, *affine[:3, :].reshape(-1))
    h[344:348] = b"n+1\0"
    payload = bytes(h) + np.asarray(data, dtype=data.dtype.newbyteorder("<")).tobytes(order="F")
    with open(path, "wb") as f:
        f.write(gzip.compress(payload, compresslevel=1) if str(path).endswith(".gz") else payload)


# ----------------------------------------------------------------------------------------------- NRRD
_NRRD_TYPES = {
    "signed char": "i1", "int8": "i1", "int8_t": "i1", "uchar": "u1", "unsigned char": "u1", "uint8": "u1", "uint8_t": "u1",
    "short": "i2", "short int": "i2", "signed short": "i2", "signed short int": "i2", "int16": "i2", "int16_t": "i2",
    "ushort": "u2", "unsigned short": "u2", "unsigned short int": "u2", "uint16": "u2", "uint16_t": "u2",
    "int": "i4", "signed int": "i4", "int32": "i4", "int32_t": "i4", "uint": "u4", "unsigned int": "u4", "uint32": "u4", "uint32_t": "u4",
    "longlong": "i8", "long long": "i8", "long long int": "i8", "signed long long": "i8", "signed long long int": "i8", "int64": "i8", "int64_t": "i8",
    "ulonglong": "u8", "unsigned long long": "u8", "unsigned long long int": "u8", "uint64": "u8", "uint64_t": "u8",
    "float": "f4", "double": "f8",
}


def _nrrd_vector(tok: str):
    tok = tok.strip()
    if tok == "none":
        return None
    if not (tok.startswith("(") and tok.endswith(")")):
        raise ValueError(f"NRRD: bad vector {tok!r}")
    return [float(x) for x in tok[1:-1].split(",")]


def read_nrrd(path) -> Tuple[np.ndarray, Dict]:
    """Returns (array shaped like ``sizes`` -- first axis fastest on disk, as pynrrd's default index order -- , header).

    Header values follow pynrrd's parsing for the fields the reference reads (vol/volume.py:868-895): ``space directions``
    float array [dim, space_dim] (NaN rows for ``none``), ``space origin`` float array, ``space`` string, ``sizes`` ints.
    """
    raw = open(path, "rb").read()
    if not raw.startswith(b"NRRD"):
        raise ValueError(f"{path}: not an NRRD file")
    end = raw.find(b"\n\n")
    sep = 2
    alt = raw.find(b"\r\n\r\n")
    if alt != -1 and (end == -1 or alt < end):
        end, sep = alt, 4
    detached_only = end == -1
    text = (raw if detached_only else raw[:end]).decode("ascii", errors="replace")
    fields: Dict[str, str] = {}
    for line in text.splitlines()[1:]:
        line = line.rstrip()
        if not line or line.startswith("#"):
            continue
        if ":=" in line:  # key/value pair
            k, v = line.split(":=", 1)
            fields[k.strip()] = v.strip()
        elif ":" in line:
            k, v = line.split(":", 1)
            fields[k.strip().lower()] = v.strip()
    for need in ("type", "dimension", "sizes", "encoding"):
        if need not in fields:
            raise ValueError(f"{path}: NRRD header lacks '{need}'")
    tname = fields["type"].lower()
    if tname not in _NRRD_TYPES:
        raise ValueError(f"{path}: unsupported NRRD type {fields['type']!r}")
    dim = int(fields["dimension"])
    sizes = [int(x) for x in fields["sizes"].split()]
    if len(sizes) != dim:
        raise ValueError(f"{path}: sizes does not match dimension")
    code = _NRRD_TYPES[tname]
    endian = {"little": "<", "big": ">"}.get(fields.get("endian", "little").lower())
    if endian is None:
        raise ValueError(f"{path}: bad endian field")
    dt = np.dtype((endian if code[1] != "1" else "|") + code)
    datafile = fields.get("data file", fields.get("datafile"))
    if datafile is not None:
        if datafile.upper().startswith("LIST") or "%" in datafile:
            raise ValueError(f"{path}: multi-file NRRD payloads are not supported")
        payload = open(os.path.join(os.path.dirname(os.path.abspath(path)), datafile), "rb").read()
    else:
        if detached_only:
            raise ValueError(f"{path}: no payload after the NRRD header")
        payload = raw[end + sep:]
    line_skip, byte_skip = int(fields.get("line skip", fields.get("lineskip", 0))), int(fields.get("byte skip", fields.get("byteskip", 0)))
    enc = fields["encoding"].lower()
    n = int(np.prod(sizes))
    if enc in ("gzip", "gz"):
        payload = zlib.decompress(payload, 16 + zlib.MAX_WBITS) if payload[:2] == b"\x1f\x8b" else zlib.decompress(payload)
    elif enc in ("bzip2", "bz2"):
        import bz2
        payload = bz2.decompress(payload)
    for _ in range(line_skip):
        payload = payload[payload.index(b"\n") + 1:]
    if enc in ("raw", "gzip", "gz", "bzip2", "bz2"):
        payload = payload[-n * dt.itemsize:] if byte_skip == -1 else payload[byte_skip:]
        if len(payload) < n * dt.itemsize:
            raise ValueError(f"{path}: NRRD payload truncated")
        arr = np.frombuffer(payload, dtype=dt, count=n)
    elif enc in ("ascii", "text", "txt"):
        arr = np.array(payload.split(), dtype=np.float64).astype(dt)
        if arr.size != n:
            raise ValueError(f"{path}: NRRD ascii payload has {arr.size} values, expected {n}")
    else:
        raise ValueError(f"{path}: unsupported NRRD encoding {fields['encoding']!r}")
    data = arr.reshape(sizes, order="F")

    header: Dict = dict(fields)
    header["type"], header["dimension"], header["sizes"], header["encoding"] = fields["type"], dim, np.array(sizes), enc
    if "space dimension" in fields:
        header["space dimension"] = int(fields["space dimension"])
    if "space directions" in fields:
        vecs = [_nrrd_vector(t) for t in re.findall(r"\([^)]*\)|none", fields["space directions"])]
        width = max(len(v) for v in vecs if v is not None)
        header["space directions"] = np.array([v if v is not None else [np.nan] * width for v in vecs], dtype=np.float64)
    if "space origin" in fields:
        header["space origin"] = np.array(_nrrd_vector(fields["space origin"]), dtype=np.float64)
    if "spacings" in fields:
        header["spacings"] = np.array([float(x) for x in fields["spacings"].split()], dtype=np.float64)
    return data, header


def write_nrrd(path, data: np.ndarray, space_directions, space_origin, space: str = "left-posterior-superior",
               encoding: str = "gzip") -> None:
    data = np.asarray(data)
    names = {"i1": "int8", "u1": "uint8", "i2": "short", "u2": "ushort", "i4": "int", "u4": "uint", "i8": "longlong", "u8": "ulonglong",
             "f4": "float", "f8": "double"}
    key = data.dtype.str[1:]
    if key not in names:
        raise ValueError(f"write_nrrd: unsupported dtype {data.dtype}")
    sd = np.asarray(space_directions, dtype=np.float64)
    so = np.asarray(space_origin, dtype=np.float64).reshape(-1)

    def vec(v):
        return "(" + ",".join(repr(float(x)) for x in v) + ")"
    lines = ["NRRD0004", f"type: {names[key]}", f"dimension: {data.ndim}", f"space: {space}", "sizes: " + " ".join(str(s) for s in data.shape),
             "space directions: " + " ".join(vec(r) for r in sd), "kinds: " + " ".join(["domain"] * data.ndim), "endian: little",
             f"encoding: {encoding}", "space origin: " + vec(so)]
    body = np.asarray(data, dtype=data.dtype.newbyteorder("<")).tobytes(order="F")
    if encoding == "gzip":
        body = gzip.compress(body, compresslevel=1)
    elif encoding != "raw":
        raise ValueError("write_nrrd: encoding must be 'raw' or 'gzip'")
    with open(path, "wb") as f:
        f.write(("\n".join(lines) + "\n\n").encode("ascii") + body)
