"""Dependency-free reader/writer for the RBC-Gym checkpoint files (HDF5 subset).

The reference writes its reset-from-checkpoint files with HDF5.jl
(`src/rbc_gym/sim/rbc_sim2D.jl:36-43,64-66`) and reads them back with
`read(h5_file, "b")[idx,:,:,:]` (`rbc_sim2D.jl:173-186`).  Neither ``h5py`` nor
``libhdf5`` is available on the GPU box, so this module parses exactly the subset
of the HDF5 file format those files use:

* superblock version 0 (8-byte offsets/lengths),
* version-1 object headers with continuation blocks,
* link messages (new style, type 0x06) or old-style symbol-table groups for the root,
* attribute messages (version 1-3) holding scalar little-endian integers/floats,
* contiguous (layout v3, class 1), unfiltered little-endian IEEE float / integer
  datasets.

HDF5 stores dimensions in C order, Julia in column-major order, so a Julia array
``(n_ep, Nx, 1, Nz)`` shows up here with shape ``(Nz, 1, Nx, n_ep)`` — the episode
index is the *fastest* varying one.

`write_checkpoint` emits the same structure (superblock v0, v1 headers, link
messages, v3 attributes, contiguous f64) so files written by this repo are readable
by the reference's HDF5.jl/h5py and by `read_file` below.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from pathlib import Path
from typing import Dict, Tuple

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class H5FormatError(ValueError):
    """Raised when a file uses an HDF5 feature outside the supported subset."""


@dataclass
class H5File:
    """Parsed contents: datasets as numpy arrays (HDF5/C order) and root attributes."""

    datasets: Dict[str, np.ndarray] = field(default_factory=dict)
    attrs: Dict[str, object] = field(default_factory=dict)


def _u(buf: bytes, off: int, n: int) -> int:
    return int.from_bytes(buf[off : off + n], "little")


def _parse_datatype(buf: bytes, off: int) -> Tuple[np.dtype, int]:
    """Datatype message -> (numpy dtype, bytes consumed). Classes 0 (int) and 1 (float)."""
    cv = buf[off]
    cls, ver = cv & 0x0F, cv >> 4
    if ver not in (1, 2, 3):
        raise H5FormatError(f"datatype version {ver}")
    bits0 = buf[off + 1]
    size = _u(buf, off + 4, 4)
    if bits0 & 1:
        raise H5FormatError("big-endian data not supported")
    if cls == 0:  # fixed point: 4 bytes of properties
        signed = bool(bits0 & 0x08)
        return np.dtype(f"<{'i' if signed else 'u'}{size}"), 8 + 4
    if cls == 1:  # floating point: 12 bytes of properties
        return np.dtype(f"<f{size}"), 8 + 12
    raise H5FormatError(f"datatype class {cls} not supported")


def _parse_dataspace(buf: bytes, off: int) -> Tuple[Tuple[int, ...], int]:
    ver = buf[off]
    rank = buf[off + 1]
    flags = buf[off + 2]
    if ver == 1:
        p = off + 8
    elif ver == 2:
        p = off + 4
    else:
        raise H5FormatError(f"dataspace version {ver}")
    dims = tuple(_u(buf, p + 8 * i, 8) for i in range(rank))
    p += 8 * rank
    if flags & 1:
        p += 8 * rank
    return dims, p - off


def _iter_messages(buf: bytes, addr: int):
    """Yield (type, flags, data_offset, size) for a version-1 object header at `addr`."""
    if buf[addr] != 1:
        raise H5FormatError(f"object header version {buf[addr]} at {addr:#x} (only v1 supported)")
    nmsg = _u(buf, addr + 2, 2)
    hsize = _u(buf, addr + 8, 4)
    blocks = [(addr + 16, hsize)]
    seen = 0
    while blocks and seen < nmsg:
        p, left = blocks.pop(0)
        end = p + left
        while p + 8 <= end and seen < nmsg:
            mtype = _u(buf, p, 2)
            msize = _u(buf, p + 2, 2)
            mflags = buf[p + 4]
            data = p + 8
            seen += 1
            if mtype == 0x10:  # continuation
                blocks.append((_u(buf, data, 8), _u(buf, data + 8, 8)))
            else:
                yield mtype, mflags, data, msize
            p = data + msize


def _pad8(n: int) -> int:
    return (n + 7) & ~7


def _parse_attribute(buf: bytes, off: int) -> Tuple[str, object]:
    ver = buf[off]
    name_sz = _u(buf, off + 2, 2)
    dt_sz = _u(buf, off + 4, 2)
    ds_sz = _u(buf, off + 6, 2)
    if ver == 1:
        p = off + 8
        name = buf[p : p + name_sz].split(b"\0")[0].decode()
        p += _pad8(name_sz)
        dtype, _ = _parse_datatype(buf, p)
        p += _pad8(dt_sz)
        dims, _ = _parse_dataspace(buf, p)
        p += _pad8(ds_sz)
    elif ver in (2, 3):
        p = off + 8 + (1 if ver == 3 else 0)
        name = buf[p : p + name_sz].split(b"\0")[0].decode()
        p += name_sz
        dtype, _ = _parse_datatype(buf, p)
        p += dt_sz
        dims, _ = _parse_dataspace(buf, p)
        p += ds_sz
    else:
        raise H5FormatError(f"attribute version {ver}")
    count = int(np.prod(dims)) if dims else 1
    val = np.frombuffer(buf, dtype=dtype, count=count, offset=p)
    if not dims:
        return name, val[0].item()
    return name, val.reshape(dims).copy()


def _parse_link(buf: bytes, off: int) -> Tuple[str, int]:
    ver, flags = buf[off], buf[off + 1]
    if ver != 1:
        raise H5FormatError(f"link message version {ver}")
    p = off + 2
    ltype = 0
    if flags & 0x08:
        ltype = buf[p]
        p += 1
    if flags & 0x04:
        p += 8
    if flags & 0x10:
        p += 1
    lsz = 1 << (flags & 3)
    nlen = _u(buf, p, lsz)
    p += lsz
    name = buf[p : p + nlen].decode()
    p += nlen
    if ltype != 0:
        raise H5FormatError("only hard links supported")
    return name, _u(buf, p, 8)


def _symbol_table_links(buf: bytes, btree: int, heap: int) -> Dict[str, int]:
    """Old-style group: walk the v1 B-tree / local heap (root groups written by h5py default)."""
    links: Dict[str, int] = {}
    if buf[heap : heap + 4] != b"HEAP":
        raise H5FormatError("bad local heap")
    heap_data = _u(buf, heap + 24, 8)

    def walk(node: int):
        if buf[node : node + 4] == b"TREE":
            level = buf[node + 5]
            used = _u(buf, node + 6, 2)
            p = node + 24
            for i in range(used):
                child = _u(buf, p + 8 + 16 * i, 8)
                walk(child) if level > 0 else walk(child)
        elif buf[node : node + 4] == b"SNOD":
            n = _u(buf, node + 6, 2)
            p = node + 8
            for i in range(n):
                e = p + 40 * i
                noff = _u(buf, e, 8)
                oaddr = _u(buf, e + 8, 8)
                s = heap_data + noff
                name = buf[s : buf.index(b"\0", s)].decode()
                links[name] = oaddr
        else:
            raise H5FormatError(f"unexpected node at {node:#x}")

    if btree != _UNDEF:
        # an empty group has a B-tree node with zero entries used
        if _u(buf, btree + 6, 2) > 0:
            walk(btree)
    return links


def _read_dataset(buf: bytes, addr: int) -> np.ndarray:
    dims = dtype = data_addr = data_size = None
    for mtype, _f, off, _sz in _iter_messages(buf, addr):
        if mtype == 0x01:
            dims, _ = _parse_dataspace(buf, off)
        elif mtype == 0x03:
            dtype, _ = _parse_datatype(buf, off)
        elif mtype == 0x08:
            ver, cls = buf[off], buf[off + 1]
            if ver != 3 or cls != 1:
                raise H5FormatError(f"layout version {ver} class {cls}: only contiguous v3 supported")
            data_addr = _u(buf, off + 2, 8)
            data_size = _u(buf, off + 10, 8)
        elif mtype == 0x0B:
            raise H5FormatError("filtered datasets not supported")
    if dims is None or dtype is None or data_addr is None:
        raise H5FormatError(f"incomplete dataset header at {addr:#x}")
    count = int(np.prod(dims))
    if count * dtype.itemsize != data_size:
        raise H5FormatError("dataset size mismatch")
    return np.frombuffer(buf, dtype=dtype, count=count, offset=data_addr).reshape(dims)


def read_file(path) -> H5File:
    """Parse a checkpoint file. Datasets are returned as read-only views in HDF5 (C) order."""
    buf = Path(path).read_bytes()
    if buf[:8] != _SIG:
        raise H5FormatError("not an HDF5 file")
    if buf[8] != 0 or buf[13] != 8 or buf[14] != 8:
        raise H5FormatError("only superblock v0 with 8-byte offsets/lengths supported")
    # superblock v0: root symbol-table entry starts at byte 56; object header address at +8
    root = _u(buf, 56 + 8, 8)
    out = H5File()
    links: Dict[str, int] = {}
    for mtype, _f, off, _sz in _iter_messages(buf, root):
        if mtype == 0x06:
            name, addr = _parse_link(buf, off)
            links[name] = addr
        elif mtype == 0x11:
            links.update(_symbol_table_links(buf, _u(buf, off, 8), _u(buf, off + 8, 8)))
        elif mtype == 0x0C:
            name, val = _parse_attribute(buf, off)
            out.attrs[name] = val
    for name, addr in links.items():
        out.datasets[name] = _read_dataset(buf, addr)
    return out


@dataclass
class Checkpoint2D:
    """Checkpoint bank in the solver's layout: arrays indexed ``[episode, z, x]`` (z=0 bottom).

    ``b`` and ``u`` have ``Nz`` rows, ``w`` has ``Nz+1`` rows (both wall faces).
    """

    b: np.ndarray
    u: np.ndarray
    w: np.ndarray
    num_episodes: int
    start_seed: int

    @property
    def shape(self) -> Tuple[int, int]:
        return self.b.shape[1], self.b.shape[2]


def load_checkpoint_2d(path) -> Checkpoint2D:
    """Read `ckpt_ra*.h5` (layout written at `rbc_sim2D.jl:36-43,64-66`) into `[ep, z, x]` float64."""
    f = read_file(path)
    arrs = {}
    for name in ("b", "u", "w"):
        a = f.datasets[name]  # (Nz or Nz+1, 1, Nx, n_ep)
        if a.ndim != 4 or a.shape[1] != 1:
            raise H5FormatError(f"dataset {name}: expected (Nz,1,Nx,n_ep), got {a.shape}")
        arrs[name] = np.ascontiguousarray(np.transpose(a[:, 0, :, :], (2, 0, 1)), dtype=np.float64)
    n = int(f.attrs.get("num_episodes", arrs["b"].shape[0]))
    if n != arrs["b"].shape[0]:
        raise H5FormatError("num_episodes attribute disagrees with dataset extent")
    return Checkpoint2D(arrs["b"], arrs["u"], arrs["w"], n, int(f.attrs.get("start_seed", 0)))


def load_checkpoint_3d(path) -> np.ndarray:
    """Read a `3D_ckpt_ra*.h5` (datasets b,u,v,w with Julia dims (n_ep, Nx, Ny, Nz|Nz+1), written at
    `rbc_sim3D.jl:57-65`) into a bank `[n_ep, 3*Nx*Ny*Nz + Nx*Ny*(Nz+1)]` float64 with fields laid out [z][y][x]."""
    f = read_file(path)
    parts = []
    for name in ("b", "u", "v", "w"):
        a = f.datasets[name]                                   # HDF5 order: (Nz|Nz+1, Ny, Nx, n_ep)
        if a.ndim != 4:
            raise H5FormatError(f"dataset {name}: expected 4 dims, got {a.shape}")
        parts.append(np.ascontiguousarray(np.moveaxis(a, 3, 0), dtype=np.float64).reshape(a.shape[3], -1))
    return np.concatenate(parts, axis=1)


# --------------------------------------------------------------------------------------
# writer (same on-disk structure as the reference files)
# --------------------------------------------------------------------------------------

def _msg(mtype: int, data: bytes, flags: int = 0) -> bytes:
    data = data + b"\0" * (_pad8(len(data)) - len(data))
    return struct.pack("<HHB3x", mtype, len(data), flags) + data


def _dt_f64() -> bytes:
    # class 1 v1, little-endian IEEE: bit fields 0x20,0x3f,0x00; size 8;
    # properties: bit offset 0, precision 64, exp loc 52, exp size 11, mant loc 0, mant size 52, bias 1023
    return bytes([0x11, 0x20, 0x3F, 0x00]) + struct.pack("<I", 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)


def _dt_i64() -> bytes:
    return bytes([0x10, 0x08, 0x00, 0x00]) + struct.pack("<I", 8) + struct.pack("<HH", 0, 64)


def _ds_simple(dims) -> bytes:
    r = len(dims)
    out = struct.pack("<BBB5x", 1, r, 1 if r else 0)
    for d in dims:
        out += struct.pack("<Q", d)
    for d in dims:  # max dims = dims
        out += struct.pack("<Q", d)
    return out


def _attr_v3_scalar_i64(name: str, value: int) -> bytes:
    nm = name.encode() + b"\0"
    dt, ds = _dt_i64(), struct.pack("<BBB1x4x", 1, 0, 0)
    body = struct.pack("<BBHHHB", 3, 0, len(nm), len(dt), len(ds), 1) + nm + dt + ds + struct.pack("<q", value)
    return body


def _object_header(messages) -> bytes:
    body = b"".join(messages)
    return struct.pack("<BxHII4x", 1, len(messages), 1, len(body)) + body


_DATASET_HEADER_BYTES = 256          # message area of a dataset header as libhdf5 allocates it (the rest is one null message)


def write_h5(path, datasets: Dict[str, np.ndarray], attrs: Dict[str, int], mtime: int | None = None) -> None:
    """Write little-endian f64 datasets (given in HDF5/C dimension order) and scalar i64 root attributes the way libhdf5
    (through HDF5.jl) lays the reference files out: superblock v0; a v1 root object header of a *new-style compact group*
    (Link Info + Group Info messages — libhdf5 looks for a symbol table unless Link Info is there — then v3 attributes and one
    hard-link message per dataset); v1 dataset headers with dataspace v1, datatype, fill value v2, contiguous layout v3, a
    modification-time message (`mtime`, seconds since the epoch; default now) and a null message filling the 256-byte message
    area; raw data from 0x800.

    Given the reference's `mtime`, the dataset headers and everything from 0x800 on are byte-identical to the file HDF5.jl wrote
    for the same arrays (`tests/test_h5lite_roundtrip.py`); the root group differs in placement only — libhdf5 grew it through
    continuation blocks scattered between the dataset headers and left the B-tree / heap of the symbol table it started with,
    here it is one contiguous block."""
    import time
    payload = {k: np.ascontiguousarray(a, dtype="<f8") for k, a in datasets.items()}
    names = list(payload)
    base, data_start = 0x60, 0x800
    stamp = int(time.time()) if mtime is None else int(mtime)

    def ds_header(arr: np.ndarray, addr: int) -> bytes:
        msgs = [
            _msg(0x01, _ds_simple(arr.shape)),
            _msg(0x03, _dt_f64(), flags=1),
            _msg(0x05, struct.pack("<BBBB", 2, 2, 2, 1), flags=1),          # fill value v2: allocate late, write if set, undefined
            _msg(0x08, struct.pack("<BBQQ", 3, 1, addr, arr.nbytes)),
            _msg(0x12, struct.pack("<B3xI", 1, stamp & 0xFFFFFFFF)),
        ]
        used = sum(len(m) for m in msgs)
        msgs.append(_msg(0x00, b"\0" * (_DATASET_HEADER_BYTES - used - 8)))
        return _object_header(msgs)

    addrs, off = {}, data_start
    for k in names:
        addrs[k] = off
        off += payload[k].nbytes
    eof = off
    group_msgs = [
        _msg(0x02, struct.pack("<BBQQ", 0, 0, _UNDEF, _UNDEF)),              # link info: compact storage (no heap, no index)
        _msg(0x0A, struct.pack("<BB", 0, 0), flags=1),                       # group info: defaults
    ] + [_msg(0x0C, _attr_v3_scalar_i64(k, int(v))) for k, v in attrs.items()]

    def link(name: str, addr: int) -> bytes:
        nm = name.encode()
        return _msg(0x06, struct.pack("<BBB", 1, 0x10, 1) + bytes([len(nm)]) + nm + struct.pack("<Q", addr))

    root_len = 16 + sum(len(m) for m in group_msgs) + sum(len(link(k, 0)) for k in names)
    p = _pad8(base + root_len)
    ds_addr, hdrs = {}, {}
    for k in names:
        ds_addr[k] = p
        hdrs[k] = ds_header(payload[k], addrs[k])
        p = _pad8(p + len(hdrs[k]))
    if p > data_start:
        raise H5FormatError("header region overflow")
    root = _object_header(group_msgs + [link(k, ds_addr[k]) for k in names])
    sb = _SIG + bytes([0, 0, 0, 0, 0, 8, 8, 0]) + struct.pack("<HHI", 4, 16, 0)
    sb += struct.pack("<QQQQ", 0, _UNDEF, eof, _UNDEF)
    sb += struct.pack("<QQII16x", 0, base, 0, 0)
    buf = bytearray(eof)
    buf[: len(sb)] = sb
    buf[base : base + len(root)] = root
    for k in names:
        buf[ds_addr[k] : ds_addr[k] + len(hdrs[k])] = hdrs[k]
        buf[addrs[k] : addrs[k] + payload[k].nbytes] = payload[k].tobytes()
    Path(path).write_bytes(bytes(buf))


def write_checkpoint(path, b: np.ndarray, u: np.ndarray, w: np.ndarray, start_seed: int, mtime: int | None = None) -> None:
    """Write 2D arrays ``[ep, z, x]`` as a reference-compatible `ckpt_ra*.h5`: datasets ``b,u,w`` with HDF5 dims
    ``(Nz|Nz+1, 1, Nx, n_ep)``, root attributes ``num_episodes`` and ``start_seed`` (`rbc_sim2D.jl:39-43`).  The data
    section and the dataset headers are what the reference writes for the same arrays (see `write_h5`)."""
    arrays = {k: np.asarray(a, dtype="<f8") for k, a in (("b", b), ("u", u), ("w", w))}
    write_h5(path, {k: np.transpose(a, (1, 2, 0))[:, None, :, :] for k, a in arrays.items()},
             {"num_episodes": arrays["b"].shape[0], "start_seed": int(start_seed)}, mtime=mtime)


def write_checkpoint_3d(path, b, u, v, w, start_seed: int, mtime: int | None = None) -> None:
    """Write 3D arrays ``[ep, z, y, x]`` as `3D_ckpt_ra*.h5` (`rbc_sim3D.jl:57-65`): HDF5 dims ``(Nz|Nz+1, Ny, Nx, n_ep)``."""
    arrays = {k: np.asarray(a, dtype="<f8") for k, a in (("b", b), ("u", u), ("v", v), ("w", w))}
    write_h5(path, {k: np.moveaxis(a, 0, 3) for k, a in arrays.items()},
             {"num_episodes": arrays["b"].shape[0], "start_seed": int(start_seed)}, mtime=mtime)
