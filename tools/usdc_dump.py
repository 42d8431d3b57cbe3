"""Minimal reader for binary USD crate files (``PXR-USDC`` 0.8): dumps prim / property specs with their authored values.

Build-container tool only (it reads ``/root/reference/.../zbot_assets/*.usd``): used once to decode the link frames, joint
frames and mass properties embedded as constants in ``zbot_lab_b200/assets/*.py``.  ``pxr`` is not installed, hence
this reader.  Supports exactly what those files use: LZ4-compressed structural sections, integer delta coding, scalar /
vector / quaternion / token / string values and small arrays.

    python tools/usdc_dump.py <file.usd> [path-substring ...]
"""
from __future__ import annotations

import struct
import sys


def lz4_block(src: bytes, out_size: int) -> bytes:
    out = bytearray()
    i, n = 0, len(src)
    while i < n:
        tok = src[i]; i += 1
        lit = tok >> 4
        if lit == 15:
            while True:
                b = src[i]; i += 1
                lit += b
                if b != 255:
                    break
        out += src[i:i + lit]; i += lit
        if i >= n:
            break
        off = src[i] | (src[i + 1] << 8); i += 2
        ml = tok & 15
        if ml == 15:
            while True:
                b = src[i]; i += 1
                ml += b
                if b != 255:
                    break
        ml += 4
        start = len(out) - off
        for k in range(ml):
            out.append(out[start + k])
    assert len(out) == out_size, (len(out), out_size)
    return bytes(out)


def fast_decompress(src: bytes, out_size: int) -> bytes:
    """pxr TfFastCompression framing: first byte = number of chunks (0 = one LZ4 block follows)."""
    nchunks = src[0]
    if nchunks == 0:
        return lz4_block(src[1:], out_size)
    out, i = b"", 1
    for _ in range(nchunks):
        (csz,) = struct.unpack_from("<i", src, i); i += 4
        chunk = src[i:i + csz]; i += csz
        out += lz4_block(chunk, min(out_size - len(out), 2147483647))
    return out


def decode_ints(buf: bytes, n: int, width: int = 4) -> list[int]:
    """pxr Usd_IntegerCompression: common delta + 2-bit codes + variable-width deltas, prefix-summed."""
    if n == 0:
        return []
    fmt_c = "<i" if width == 4 else "<q"
    (common,) = struct.unpack_from(fmt_c, buf, 0)
    codes_off = width
    vals_off = codes_off + (n * 2 + 7) // 8
    small = [(1, "<b"), (2, "<h"), (4, "<i")] if width == 4 else [(2, "<h"), (4, "<i"), (8, "<q")]
    out, prev, vi = [], 0, vals_off
    for k in range(n):
        code = (buf[codes_off + k // 4] >> ((k % 4) * 2)) & 3
        if code == 0:
            d = common
        else:
            sz, f = small[code - 1]
            (d,) = struct.unpack_from(f, buf, vi); vi += sz
        prev += d
        out.append(prev)
    return out


class Crate:
    def __init__(self, path: str):
        self.d = open(path, "rb").read()
        d = self.d
        assert d[:8] == b"PXR-USDC", d[:8]
        self.version = tuple(d[8:11])
        (toc,) = struct.unpack_from("<q", d, 16)
        (nsec,) = struct.unpack_from("<Q", d, toc)
        self.sections = {}
        for k in range(nsec):
            name, start, size = struct.unpack_from("<16sqq", d, toc + 8 + 32 * k)
            self.sections[name.split(b"\0")[0].decode()] = (start, size)
        self._tokens()
        self._strings()
        self._fields()
        self._fieldsets()
        self._paths()
        self._specs()

    def _compressed_ints(self, off: int, n: int, width: int = 4):
        (csz,) = struct.unpack_from("<Q", self.d, off); off += 8
        raw = self.d[off:off + csz]
        enc_size = width + (n * 2 + 7) // 8 + n * width if n else 0
        buf = fast_decompress_any(raw, enc_size)
        return decode_ints(buf, n, width), off + csz

    def _tokens(self):
        o, _ = self.sections["TOKENS"]
        n, usz, csz = struct.unpack_from("<QQQ", self.d, o)
        raw = fast_decompress(self.d[o + 24:o + 24 + csz], usz)
        self.tokens = [t.decode("utf8", "replace") for t in raw.split(b"\0")[:n]]

    def _strings(self):
        o, _ = self.sections["STRINGS"]
        (n,) = struct.unpack_from("<Q", self.d, o)
        self.strings = list(struct.unpack_from(f"<{n}I", self.d, o + 8))

    def _fields(self):
        o, _ = self.sections["FIELDS"]
        (n,) = struct.unpack_from("<Q", self.d, o)
        tok, o2 = self._compressed_ints(o + 8, n)
        (csz,) = struct.unpack_from("<Q", self.d, o2)
        reps = fast_decompress(self.d[o2 + 8:o2 + 8 + csz], n * 8)
        self.fields = [(self.tokens[tok[k]], struct.unpack_from("<Q", reps, 8 * k)[0]) for k in range(n)]

    def _fieldsets(self):
        o, _ = self.sections["FIELDSETS"]
        (n,) = struct.unpack_from("<Q", self.d, o)
        vals, _ = self._compressed_ints(o + 8, n)
        self.fieldsets = [v & 0xFFFFFFFF for v in vals]

    def _paths(self):
        o, _ = self.sections["PATHS"]
        (npaths,) = struct.unpack_from("<Q", self.d, o)
        (nenc,) = struct.unpack_from("<Q", self.d, o + 8)
        pidx, o2 = self._compressed_ints(o + 16, nenc)
        etok, o2 = self._compressed_ints(o2, nenc)
        jumps, o2 = self._compressed_ints(o2, nenc)
        self.paths = [None] * npaths
        stack = [(0, "")]
        while stack:
            cur, parent = stack.pop()
            while True:
                this = cur
                cur += 1
                if parent == "":
                    path = "/"
                else:
                    t = etok[this]
                    prop = t < 0
                    name = self.tokens[abs(t)]
                    path = parent + ("." if prop else ("" if parent == "/" else "/")) + name
                self.paths[pidx[this] & 0xFFFFFFFF] = path
                j = jumps[this]
                has_child = j > 0 or j == -1
                has_sibling = j >= 0
                if has_child:
                    if has_sibling:
                        stack.append((this + j, parent))
                    parent = path
                    continue
                if has_sibling:
                    continue
                break

    def _specs(self):
        o, _ = self.sections["SPECS"]
        (n,) = struct.unpack_from("<Q", self.d, o)
        pi, o2 = self._compressed_ints(o + 8, n)
        fs, o2 = self._compressed_ints(o2, n)
        st, o2 = self._compressed_ints(o2, n)
        self.specs = [(self.paths[pi[k] & 0xFFFFFFFF], fs[k] & 0xFFFFFFFF, st[k]) for k in range(n)]

    # ---- values -------------------------------------------------------------------------------------
    def value(self, rep: int):
        is_array = bool(rep >> 63 & 1)
        inlined = bool(rep >> 62 & 1)
        compressed = bool(rep >> 61 & 1)
        ty = (rep >> 48) & 0xFF
        payload = rep & ((1 << 48) - 1)
        d = self.d
        scalar = {1: ("<?", 1), 2: ("<B", 1), 3: ("<i", 4), 4: ("<I", 4), 5: ("<q", 8), 6: ("<Q", 8), 7: ("<e", 2),
                  8: ("<f", 4), 9: ("<d", 8)}
        vec = {16: ("d", 4), 17: ("f", 4), 18: ("e", 4), 19: ("d", 2), 20: ("f", 2), 21: ("e", 2), 22: ("i", 2),
               23: ("d", 3), 24: ("f", 3), 25: ("e", 3), 26: ("i", 3), 27: ("d", 4), 28: ("f", 4), 29: ("e", 4),
               30: ("i", 4), 13: ("d", 4), 14: ("d", 9), 15: ("d", 16)}
        if not is_array:
            if ty in scalar:
                f, sz = scalar[ty]
                if inlined:
                    if ty == 9:     # double stored as float when exactly representable
                        return struct.unpack("<f", struct.pack("<I", payload & 0xFFFFFFFF))[0]
                    if ty in (5, 6):
                        return struct.unpack("<i", struct.pack("<I", payload & 0xFFFFFFFF))[0]
                    return struct.unpack_from(f, struct.pack("<Q", payload))[0]
                return struct.unpack_from(f, d, payload)[0]
            if ty == 10:
                return self.tokens[self.strings[payload]]
            if ty in (11, 12):
                return self.tokens[payload]
            if ty in vec:
                c, k = vec[ty]
                if inlined:     # small-integer components packed as int8 (matrices: the diagonal)
                    b = struct.pack("<Q", payload)
                    if ty in (13, 14, 15):
                        return ("diag", [struct.unpack_from("<b", b, i)[0] for i in range({13: 2, 14: 3, 15: 4}[ty])])
                    return tuple(float(struct.unpack_from("<b", b, i)[0]) for i in range(k))
                v = struct.unpack_from(f"<{k}{c}", d, payload)
                if ty in (16, 17, 18):   # GfQuat memory order: imaginary (x, y, z), real  ->  report (w, x, y, z)
                    return (v[3], v[0], v[1], v[2])
                return v
            if ty == 42:
                return ("def", "over", "class")[payload] if payload < 3 else payload
            if ty == 44:
                return ("varying", "uniform")[payload] if payload < 2 else payload
            if ty == 41:    # token vector
                (n,) = struct.unpack_from("<Q", d, payload)
                return [self.tokens[i] for i in struct.unpack_from(f"<{n}I", d, payload + 8)]
            if ty == 32:    # token list op
                return self._listop(payload, lambda o, n: [self.tokens[i] for i in struct.unpack_from(f"<{n}I", d, o)], 4)
            if ty == 34:    # path list op
                return self._listop(payload, lambda o, n: [self.paths[i] for i in struct.unpack_from(f"<{n}I", d, o)], 4)
            return f"<type {ty} @{payload}>"
        # arrays
        if payload == 0:
            return []
        (n,) = struct.unpack_from("<Q", d, payload)
        o = payload + 8
        if ty in scalar:
            f, sz = scalar[ty]
            if compressed and ty in (3, 4, 5, 6):
                (csz,) = struct.unpack_from("<Q", d, o)
                w = 4 if ty in (3, 4) else 8
                buf = fast_decompress_any(d[o + 8:o + 8 + csz], w + (n * 2 + 7) // 8 + n * w)
                return decode_ints(buf, n, w)
            if compressed:
                return f"<compressed float array n={n}>"
            return list(struct.unpack_from(f"<{n}{f[1]}", d, o))
        if ty == 11:
            return [self.tokens[i] for i in struct.unpack_from(f"<{n}I", d, o)]
        if ty in vec:
            c, k = vec[ty]
            flat = struct.unpack_from(f"<{n * k}{c}", d, o)
            if n > 8:
                return f"<{n} x vec{k}{c}>"
            return [flat[i * k:(i + 1) * k] for i in range(n)]
        return f"<array type {ty} n={n}>"

    def _listop(self, off, read, width):
        d = self.d
        bits = d[off]; off += 1
        out = {}
        for name, mask in (("explicit", 2), ("added", 4), ("prepended", 32), ("appended", 64), ("deleted", 8), ("ordered", 16)):
            if bits & mask:
                (n,) = struct.unpack_from("<Q", d, off); off += 8
                out[name] = read(off, n); off += n * width
        return out

    def spec_fields(self, fs_index: int):
        out = {}
        k = fs_index
        while self.fieldsets[k] != 0xFFFFFFFF:
            name, rep = self.fields[self.fieldsets[k]]
            out[name] = rep
            k += 1
        return out


def fast_decompress_any(raw: bytes, max_size: int) -> bytes:
    """Decompress when only an upper bound of the output size is known (integer sections)."""
    nchunks = raw[0]
    assert nchunks == 0, "multi-chunk integer sections not needed for these files"
    return _lz4_unbounded(raw[1:])


def _lz4_unbounded(src: bytes) -> bytes:
    out = bytearray()
    i, n = 0, len(src)
    while i < n:
        tok = src[i]; i += 1
        lit = tok >> 4
        if lit == 15:
            while True:
                b = src[i]; i += 1
                lit += b
                if b != 255:
                    break
        out += src[i:i + lit]; i += lit
        if i >= n:
            break
        off = src[i] | (src[i + 1] << 8); i += 2
        ml = tok & 15
        if ml == 15:
            while True:
                b = src[i]; i += 1
                ml += b
                if b != 255:
                    break
        ml += 4
        start = len(out) - off
        for k in range(ml):
            out.append(out[start + k])
    return bytes(out)


def main(argv):
    c = Crate(argv[1])
    filt = argv[2:]
    skip = {"points", "normals", "faceVertexIndices", "faceVertexCounts", "primvars:st", "primvars:normals"}
    for path, fs, st in c.specs:
        if filt and not any(f in path for f in filt):
            continue
        leaf = path.rsplit(".", 1)[-1] if "." in path else ""
        if leaf in skip or "Looks" in path or "visuals" in path:
            continue
        fields = c.spec_fields(fs)
        vals = {}
        for name, rep in fields.items():
            if name in ("typeName", "default", "specifier", "apiSchemas", "targetPaths", "variability", "custom"):
                try:
                    vals[name] = c.value(rep)
                except Exception as e:  # noqa: BLE001
                    vals[name] = f"<err {e}>"
        if "." in path and "default" not in vals and "targetPaths" not in vals:
            continue
        print(path, vals)


if __name__ == "__main__":
    main(sys.argv)
