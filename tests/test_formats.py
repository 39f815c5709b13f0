"""NIfTI-1 / NRRD loaders (reference vol/volume.py:581-696, 848-895).

nibabel and pynrrd are not available offline, so the fixtures are assembled here byte by byte from the two format
specifications (independently of the package's own writers) and the readers are checked against the known content;
the writers are then checked by round trip.
"""
import gzip
import struct

import numpy as np
import pytest

from deepdrr_b200 import formats
from deepdrr_b200.vol import Volume


def _nifti_bytes(vox, endian="<", code=4, slope=1.0, inter=0.0, sform=None, quatern=None, pixdim=(1, 1, 1, 1), vox_offset=352.0, dim4=False):
    h = bytearray(int(vox_offset))
    e = endian
    struct.pack_into(e + "i", h, 0, 348)
    if dim4:
        struct.pack_into(e + "8h", h, 40, 4, *vox.shape, 1, 1, 1, 1)
    else:
        struct.pack_into(e + "8h", h, 40, 3, *vox.shape, 1, 1, 1, 1)
    struct.pack_into(e + "2h", h, 70, code, vox.dtype.itemsize * 8)
    struct.pack_into(e + "8f", h, 76, *pixdim, 0, 0, 0, 0)
    struct.pack_into(e + "3f", h, 108, vox_offset, slope, inter)
    h[123] = 2
    if sform is not None:
        struct.pack_into(e + "h", h, 254, 1)
        struct.pack_into(e + "12f", h, 280, *np.asarray(sform, dtype=np.float64)[:3].reshape(-1))
    if quatern is not None:
        struct.pack_into(e + "h", h, 252, 1)
        struct.pack_into(e + "6f", h, 256, *quatern)
    h[344:348] = b"n+1\0"
    return bytes(h) + vox.astype(vox.dtype.newbyteorder(e)).tobytes(order="F")


def test_nifti_int16_scaled_sform_gz(tmp_path):
    rng = np.random.default_rng(0)
    vox = rng.integers(-1000, 2000, size=(5, 6, 7)).astype(np.int16)
    aff = np.array([[0.8, 0, 0, -10], [0, 0, 1.5, 3], [0, -0.9, 0, 7.5], [0, 0, 0, 1]])
    p = tmp_path / "a.nii.gz"
    p.write_bytes(gzip.compress(_nifti_bytes(vox, slope=2.0, inter=-1024.0, sform=aff, vox_offset=400.0)))
    data, affine, hdr = formats.read_nifti(p)
    assert data.dtype == np.float64 and data.shape == (5, 6, 7)
    assert np.array_equal(data, vox.astype(np.float64) * 2.0 - 1024.0)
    assert np.allclose(affine, aff, atol=1e-6) and hdr["xyz_units"] == "mm"
    assert data[1, 2, 3] == vox[1, 2, 3] * 2.0 - 1024.0     # first index fastest on disk


def test_nifti_big_endian_float_qform(tmp_path):
    rng = np.random.default_rng(1)
    vox = rng.normal(size=(4, 3, 2)).astype(np.float32)
    # 90 degrees about z: quaternion (a, b, c, d) = (cos 45, 0, 0, sin 45); qfac = -1 flips the third axis
    s = np.sqrt(0.5)
    p = tmp_path / "b.nii"
    p.write_bytes(_nifti_bytes(vox, endian=">", code=16, quatern=(0.0, 0.0, s, 11.0, 12.0, 13.0), pixdim=(-1.0, 2.0, 3.0, 4.0), dim4=True))
    data, affine, hdr = formats.read_nifti(p)
    assert np.array_equal(data, vox.astype(np.float64)) and hdr["endian"] == ">"
    expect = np.array([[0, -3.0, 0, 11.0], [2.0, 0, 0, 12.0], [0, 0, -4.0, 13.0], [0, 0, 0, 1]])
    assert np.allclose(affine, expect, atol=1e-6)


def test_nifti_without_forms_and_errors(tmp_path):
    vox = np.arange(24, dtype=np.uint8).reshape(2, 3, 4)
    p = tmp_path / "c.nii"
    p.write_bytes(_nifti_bytes(vox, code=2, pixdim=(1.0, 0.5, 0.5, 2.0)))
    data, affine, _ = formats.read_nifti(p)
    assert np.array_equal(data, vox)
    assert np.allclose(np.diag(affine), [-0.5, 0.5, 2.0, 1.0]) and np.allclose(affine[:3, 3], [0.25, -0.5, -3.0])
    bad = bytearray(p.read_bytes()); bad[344:348] = b"xxxx"
    (tmp_path / "bad.nii").write_bytes(bytes(bad))
    with pytest.raises(ValueError):
        formats.read_nifti(tmp_path / "bad.nii")
    (tmp_path / "short.nii").write_bytes(p.read_bytes()[:360])
    with pytest.raises(ValueError):
        formats.read_nifti(tmp_path / "short.nii")


def test_nifti_writer_round_trip_and_volume(tmp_path):
    rng = np.random.default_rng(2)
    hu = rng.uniform(-1000, 1500, size=(6, 5, 4)).astype(np.float32)
    aff = np.array([[-0.7, 0, 0, 20], [0, -0.7, 0, 30], [0, 0, 1.25, -40], [0, 0, 0, 1]])
    p = tmp_path / "ct.nii.gz"
    formats.write_nifti(p, hu, aff)
    data, affine, _ = formats.read_nifti(p)
    assert np.array_equal(data, hu.astype(np.float64)) and np.allclose(affine, aff, atol=1e-5)
    v = Volume.from_nifti(p)
    w = Volume.from_hu(hu.astype(np.float64))
    assert np.array_equal(v.data, w.data) and np.array_equal(v.materials[1], w.materials[1]) and v.materials[0] == w.materials[0]
    assert v.anatomical_coordinate_system == "RAS" and np.allclose(v.anatomical_from_IJK.data, aff, atol=1e-5)
    assert np.allclose(v.spacing, [0.7, 0.7, 1.25], atol=1e-6)
    # segmentation files: label selection and binarisation (reference vol/volume.py:643-660)
    lab = rng.integers(0, 4, size=(6, 5, 4)).astype(np.uint8)
    q = tmp_path / "seg.nii"
    formats.write_nifti(q, lab, aff)
    s = Volume.from_nifti(q, segmentation=True, label=[2, 3], binarize=True)
    assert s.materials[0] == {"bone": 0} and np.array_equal(s.data, np.isin(lab, [2, 3]).astype(np.float32))
    s1 = Volume.from_nifti(q, segmentation=True)
    assert np.array_equal(s1.data, lab.astype(np.float32))
    # explicit masks, one of them by path
    m = Volume.from_nifti(p, materials={"air": hu <= 0, "bone": str(q)})
    assert list(m.materials[0]) == ["air", "bone"]
    assert np.array_equal(m.materials[1], np.where(lab > 0, 1, 0).astype(np.uint16))
    with pytest.raises(NotImplementedError):
        Volume.from_nifti(p, use_thresholding=False)


_NRRD_HEAD = """NRRD0004
# Complete NRRD file format specification at:
# http://teem.sourceforge.net/nrrd/format.html
type: short
dimension: 3
space: left-posterior-superior
sizes: 4 3 2
space directions: (0.5,0,0) (0,0.5,0) (0,0,2.5)
kinds: domain domain domain
endian: {endian}
encoding: {enc}
space origin: (-10.5,20,30.25)
"""


def test_nrrd_raw_gzip_ascii_detached(tmp_path):
    vox = (np.arange(24, dtype=np.int16) * 37 - 400).reshape(4, 3, 2, order="F")
    raw_le = vox.astype("<i2").tobytes(order="F")
    cases = {
        "raw.nrrd": _NRRD_HEAD.format(endian="little", enc="raw").encode() + b"\n" + raw_le,
        "big.nrrd": _NRRD_HEAD.format(endian="big", enc="raw").encode() + b"\n" + vox.astype(">i2").tobytes(order="F"),
        "gz.nrrd": _NRRD_HEAD.format(endian="little", enc="gzip").encode() + b"\n" + gzip.compress(raw_le),
        "txt.nrrd": _NRRD_HEAD.format(endian="little", enc="ascii").encode() + b"\n" + " ".join(str(int(x)) for x in vox.reshape(-1, order="F")).encode() + b"\n",
        "crlf.nrrd": _NRRD_HEAD.format(endian="little", enc="raw").replace("\n", "\r\n").encode() + b"\r\n" + raw_le,
        "det.nhdr": (_NRRD_HEAD.format(endian="little", enc="raw") + "data file: det.raw\n").encode(),
    }
    (tmp_path / "det.raw").write_bytes(raw_le)
    for name, blob in cases.items():
        (tmp_path / name).write_bytes(blob)
        data, header = formats.read_nrrd(tmp_path / name)
        assert data.shape == (4, 3, 2) and np.array_equal(data, vox), name
        assert np.array_equal(header["space directions"], np.diag([0.5, 0.5, 2.5]))
        assert np.array_equal(header["space origin"], [-10.5, 20, 30.25]) and header["space"] == "left-posterior-superior"
    v = Volume.from_nrrd(tmp_path / "gz.nrrd")
    w = Volume.from_hu(vox)
    assert np.array_equal(v.data, w.data) and np.array_equal(v.materials[1], w.materials[1])
    assert v.anatomical_coordinate_system == "LPS"
    expect = np.array([[0.5, 0, 0, -10.5], [0, 0.5, 0, 20], [0, 0, 2.5, 30.25], [0, 0, 0, 1]])
    assert np.array_equal(v.anatomical_from_IJK.data, expect)
    with pytest.raises(ValueError):
        (tmp_path / "no.nrrd").write_bytes(b"hello")
        formats.read_nrrd(tmp_path / "no.nrrd")


def test_nrrd_writer_round_trip_and_row_convention(tmp_path):
    rng = np.random.default_rng(3)
    hu = rng.uniform(-1000, 1000, size=(3, 4, 5)).astype(np.float32)
    dirs = np.array([[0.0, 0.6, 0.0], [0.7, 0.0, 0.0], [0.0, 0.0, 1.1]])     # axis 0 runs along y, axis 1 along x
    for enc in ("raw", "gzip"):
        p = tmp_path / f"w_{enc}.nrrd"
        formats.write_nrrd(p, hu, dirs, (1.0, 2.0, 3.0), space="right-anterior-superior", encoding=enc)
        data, header = formats.read_nrrd(p)
        assert np.array_equal(data, hu) and np.array_equal(header["space directions"], dirs)
    v = Volume.from_nrrd(p)
    # the reference stacks the direction vectors as ROWS (vol/volume.py:870-880): kept as is
    assert np.array_equal(v.anatomical_from_IJK.data[:3, :3], dirs) and v.anatomical_coordinate_system == "RAS"
