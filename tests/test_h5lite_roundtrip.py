"""The checkpoint writer against a file HDF5.jl wrote (`data/checkpoints/train/ckpt_ra100000.h5`, one of the reference's own
`ckpt_ra*.h5`, `rbc_sim2D.jl:36-43`): re-serialise it and compare at the byte level.  No libhdf5 exists in this image, so the
claim is structural: same size; every byte from the first dataset on identical; the three dataset object headers identical
byte for byte (given the same modification time); the root group carries the same messages — libhdf5 just placed them in
continuation blocks between the dataset headers and left the B-tree and heap of the symbol table the group started with."""
import struct
from pathlib import Path

import numpy as np

from rbc_gym_b200 import h5lite as H

SRC = Path(__file__).resolve().parents[1] / "data/checkpoints/train/ckpt_ra100000.h5"


def _messages(buf, addr):
    """All messages of the v1 object header at `addr` except continuations: (type, flags, payload bytes)."""
    return [(t, f, bytes(buf[off:off + sz])) for t, f, off, sz in H._iter_messages(buf, addr)]


def _root_and_links(buf):
    root = H._u(buf, 64, 8)
    links = dict(H._parse_link(buf, off) for t, _f, off, _sz in H._iter_messages(buf, root) if t == 0x06)
    return root, links


def test_rewritten_reference_checkpoint_is_byte_identical_outside_the_root_group(tmp_path):
    ref = SRC.read_bytes()
    c = H.load_checkpoint_2d(SRC)
    root, links = _root_and_links(ref)
    stamp = next(struct.unpack("<I", m[4:8])[0] for t, _f, m in _messages(ref, links["b"]) if t == 0x12)
    out = tmp_path / "again.h5"
    H.write_checkpoint(out, c.b, c.u, c.w, start_seed=int(H.read_file(SRC).attrs["start_seed"]), mtime=stamp)
    new = out.read_bytes()
    assert len(new) == len(ref)
    assert new[:56] == ref[:56]                                            # signature, versions, sizes, base / end-of-file addresses
    assert new[0x800:] == ref[0x800:]                                      # all raw data, at the same addresses
    nroot, nlinks = _root_and_links(new)
    assert set(nlinks) == set(links) == {"b", "u", "w"}
    for k in "buw":                                                       # dataset headers: prefix + 256-byte message area
        assert new[nlinks[k]:nlinks[k] + 16 + 256] == ref[links[k]:links[k] + 16 + 256], k
    # root group: same messages up to order and the dataset header addresses inside the link messages
    def canon(buf, addr):
        out = []
        for t, f, m in _messages(buf, addr):
            if t == 0x00:
                continue
            if t == 0x06:
                m = m[:5]                                                  # drop the target address (placement) and the padding
            out.append((t, f, m))
        return sorted(out)
    assert canon(new, nroot) == canon(ref, root)
    assert {t for t, _f, _m in _messages(ref, root)} == {0x02, 0x0A, 0x0C, 0x06}
    # and the reader sees the same thing in both
    a, b = H.read_file(SRC), H.read_file(out)
    assert a.attrs == b.attrs and all(np.array_equal(a.datasets[k], b.datasets[k]) for k in "buw")


def test_3d_checkpoint_writer_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    b, u, v = (rng.standard_normal((3, 4, 6, 8)) for _ in range(3))
    w = rng.standard_normal((3, 5, 6, 8))
    H.write_checkpoint_3d(tmp_path / "c3.h5", b, u, v, w, start_seed=7)
    f = H.read_file(tmp_path / "c3.h5")
    assert f.attrs == {"num_episodes": 3, "start_seed": 7}
    assert f.datasets["w"].shape == (5, 6, 8, 3)                           # HDF5 (C) order of Julia's (n_ep, Nx, Ny, Nz)
    assert np.array_equal(np.moveaxis(f.datasets["v"], 3, 0), v)
