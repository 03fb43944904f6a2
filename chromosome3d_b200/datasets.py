"""Input data: the reference's Hi-C IF matrices and shipped models (data, not code), plus the
synthetic higher-resolution matrices of SURVEY §8(d).

The 45 `input/chr*_matrix.txt` files and the shipped `output_models/*.pdb` travel as two xz
archives under data/ and are unpacked on first use into data/_cache/ (git-ignored).
"""
import os
import tarfile

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_DATA = os.path.join(_ROOT, "data")
_CACHE = os.path.join(_DATA, "_cache")


def _unpack(archive, marker):
    """Unpacks once, atomically: several processes (the ranks of a torchrun launch) may call this at the same time on a fresh
    checkout, so the archive is extracted into a private temporary directory and renamed into place; a rank that loses the
    rename race discards its copy.  Nobody ever sees a half-extracted directory."""
    dst = os.path.join(_CACHE, marker)
    if not os.path.isdir(dst):
        import shutil
        import tempfile
        os.makedirs(_CACHE, exist_ok=True)
        tmp = tempfile.mkdtemp(prefix=".unpack_", dir=_CACHE)
        try:
            with tarfile.open(os.path.join(_DATA, archive), "r:xz") as tf:
                tf.extractall(tmp)
            try:
                os.rename(os.path.join(tmp, marker), dst)
            except OSError:
                if not os.path.isdir(dst):
                    raise
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    return dst


def input_dir():
    """Directory holding chr*_matrix.txt (byte-identical to the reference's input/)."""
    return _unpack("hic_inputs.tar.xz", "input")


def shipped_models_dir():
    return _unpack("shipped_models.tar.xz", "output_models")


def input_path(name):
    """name: 'chr21_1mb' or a file name."""
    if not name.endswith(".txt"):
        name = name + "_matrix.txt"
    return os.path.join(input_dir(), name)


def list_inputs():
    return sorted(f[:-len("_matrix.txt")] for f in os.listdir(input_dir()) if f.endswith("_matrix.txt"))


def load_if_matrix(path):
    """Parse an IF matrix the way IF2dist_new does (C3D:116-129): one row per line, tokens split
    on whitespace (this eats the CR of the CRLF line ends); L = token count of the first row
    (calc_len_IF, C3D:164-179).  Returns a C-contiguous float64 (L, L) array."""
    rows = []
    with open(path, "r", newline="") as fh:
        for line in fh:
            toks = line.split()
            if not toks:
                continue
            rows.append([float(t) for t in toks])
    L = len(rows[0])
    a = np.zeros((L, L), dtype=np.float64)
    for i, r in enumerate(rows[:L]):
        a[i, :min(L, len(r))] = r[:L]
    return np.ascontiguousarray(a)


def read_ca_pdb(path):
    """CA coordinates of a PDB as the text fields give them (C3D:674-691 columns), in file order."""
    xyz = []
    with open(path, "r", newline="") as fh:
        for line in fh:
            if not line.startswith("ATOM"):
                continue
            if line[12:16].strip() != "CA":
                continue
            xyz.append((float(line[30:38]), float(line[38:46]), float(line[46:54])))
    return np.ascontiguousarray(np.array(xyz, dtype=np.float64))


def philox4x32(ctr, key):
    """Philox4x32-10, vectorised over the leading axis of ctr (n,4) uint32; key (2,) uint32."""
    c = np.array(ctr, dtype=np.uint64).reshape(-1, 4).copy()
    k0, k1 = np.uint64(int(key[0])), np.uint64(int(key[1]))
    m0, m1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = m0 * c[:, 0]
        p1 = m1 * c[:, 2]
        n0 = ((p1 >> np.uint64(32)) ^ c[:, 1] ^ k0) & mask
        n1 = p1 & mask
        n2 = ((p0 >> np.uint64(32)) ^ c[:, 3] ^ k1) & mask
        n3 = p0 & mask
        c = np.stack([n0, n1, n2, n3], axis=1)
        k0 = (k0 + np.uint64(0x9E3779B9)) & mask
        k1 = (k1 + np.uint64(0xBB67AE85)) & mask
    return c.astype(np.uint32)


def synthetic_if(N, seed, A=6.8e4, slope=-0.95, sigma=0.6, zero_frac=5e-4, dtype=np.float64, block=2048):
    """Synthetic IF matrix of SURVEY §8(d): IF[i][j] = A |i-j|^slope exp(sigma z_ij), z ~ N(0,1)
    from Philox(seed, i*N+j) for i<j; diagonal 4A exp(0.3 z_ii); a zero_frac fraction of
    off-diagonal pairs set to 0.  Symmetric.  Built block-wise so N=25k stays within memory."""
    out = np.zeros((N, N), dtype=dtype)
    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    for i0 in range(0, N, block):
        i1 = min(N, i0 + block)
        for j0 in range(i0, N, block):
            j1 = min(N, j0 + block)
            ii, jj = np.meshgrid(np.arange(i0, i1, dtype=np.uint64), np.arange(j0, j1, dtype=np.uint64), indexing="ij")
            lin = (ii * np.uint64(N) + jj).ravel()
            ctr = np.zeros((lin.size, 4), dtype=np.uint64)
            ctr[:, 0] = lin & np.uint64(0xFFFFFFFF)
            ctr[:, 1] = lin >> np.uint64(32)
            r = philox4x32(ctr, key).astype(np.float64)
            u1 = (r[:, 0] + 0.5) / 4294967296.0
            u2 = (r[:, 1] + 0.5) / 4294967296.0
            z = np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)
            u3 = (r[:, 2] + 0.5) / 4294967296.0
            sep = np.abs(ii.astype(np.int64) - jj.astype(np.int64)).ravel().astype(np.float64)
            with np.errstate(divide="ignore"):
                v = np.where(sep > 0, A * np.power(np.maximum(sep, 1.0), slope) * np.exp(sigma * z),
                             4.0 * A * np.exp(0.3 * z))
            v = np.where((sep > 0) & (u3 < zero_frac), 0.0, v)
            blk = v.reshape(i1 - i0, j1 - j0)
            if i0 == j0:
                blk = np.triu(blk) + np.triu(blk, 1).T
                out[i0:i1, j0:j1] = blk
            else:
                out[i0:i1, j0:j1] = blk
                out[j0:j1, i0:i1] = blk.T
    return out
