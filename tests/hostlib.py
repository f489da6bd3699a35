"""ctypes access to the host layer (dashing_b200/host/libdashing_b200_host.so) and helpers shared by the host / CLI tests."""
import ctypes as C
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST_SO = os.path.join(ROOT, "dashing_b200", "host", "libdashing_b200_host.so")
CLI = os.path.join(ROOT, "dashing_b200", "host", "dashing_b200")


def load():
    l = C.CDLL(HOST_SO)
    l.db200h_make_fname.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_uint64]
    l.db200h_write_hll.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double]
    l.db200h_read_hll.argtypes = [C.c_char_p, C.c_void_p, C.c_uint64, C.c_void_p, C.POINTER(C.c_double)]
    l.db200h_format_symmetric.restype = C.c_uint64
    l.db200h_format_symmetric.argtypes = [C.c_char_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int, C.c_char_p, C.c_uint64]
    l.db200h_format_sizes.restype = C.c_uint64
    l.db200h_format_sizes.argtypes = [C.c_char_p, C.c_uint64, C.c_void_p, C.c_char_p, C.c_uint64]
    l.db200h_format_rect_row.restype = C.c_uint64
    l.db200h_format_rect_row.argtypes = [C.c_char_p, C.c_void_p, C.c_uint64, C.c_char_p, C.c_uint64]
    l.db200h_format_neighbors.restype = C.c_uint64
    l.db200h_format_neighbors.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_char_p, C.c_uint64]
    l.db200h_read_records.restype = C.c_int64
    l.db200h_read_records.argtypes = [C.c_char_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
    l.db200h_record_names.restype = C.c_int64
    l.db200h_record_names.argtypes = [C.c_char_p, C.c_char_p, C.c_uint64]
    l.db200h_get_paths.restype = C.c_int64
    l.db200h_get_paths.argtypes = [C.c_char_p, C.c_char_p, C.c_uint64]
    l.db200h_slurp.restype = C.c_int64
    l.db200h_slurp.argtypes = [C.c_char_p, C.c_void_p, C.c_uint64]
    l.db200h_file_capacity.restype = C.c_uint64
    l.db200h_file_capacity.argtypes = [C.c_char_p]
    return l


def make_fname(l, path, p, k, suffix="", prefix=""):
    buf = C.create_string_buffer(4096)
    assert l.db200h_make_fname(path.encode(), p, k, k, k, b"", suffix.encode(), prefix.encode(), buf, 4096) == 0
    return buf.value.decode()


def format_symmetric(l, names, packed, fmt, lower=None):
    packed = np.ascontiguousarray(packed, dtype=np.float32)
    lo = None if lower is None else np.ascontiguousarray(lower, dtype=np.float32)
    nm = "\n".join(names).encode() + b"\n"
    cap = 1 << 20
    buf = C.create_string_buffer(cap)
    n = l.db200h_format_symmetric(nm, len(names), packed.ctypes.data, None if lo is None else lo.ctypes.data, fmt, buf, cap)
    assert n <= cap
    return buf.raw[:n]


def format_sizes(l, names, card):
    card = np.ascontiguousarray(card, dtype=np.float64)
    cap = 1 << 20
    buf = C.create_string_buffer(cap)
    n = l.db200h_format_sizes("\n".join(names).encode() + b"\n", len(names), card.ctypes.data, buf, cap)
    return buf.raw[:n]


def format_neighbors(l, names, qoffset, nb, fmt):
    """nb: structured array [rows][nn] of (value f32, index u32)."""
    nb = np.ascontiguousarray(nb)
    rows, nn = nb.shape
    cap = 1 << 20
    buf = C.create_string_buffer(cap)
    n = l.db200h_format_neighbors("\n".join(names).encode() + b"\n", len(names), qoffset, nb.ctypes.data, rows, nn, fmt, buf, cap)
    assert n <= cap
    return buf.raw[:n]


def format_rect_row(l, qname, row):
    row = np.ascontiguousarray(row, dtype=np.float32)
    buf = C.create_string_buffer(1 << 16)
    n = l.db200h_format_rect_row(qname.encode(), row.ctypes.data, row.size, buf, 1 << 16)
    return buf.raw[:n]


def read_records(l, path, cap=1 << 22):
    bases = np.zeros(cap, dtype=np.uint8)
    offs = np.zeros(100000, dtype=np.uint64)
    n = l.db200h_read_records(os.fsencode(path), bases.ctypes.data, cap, offs.ctypes.data, offs.size - 1)
    assert n >= 0, n
    return [bases[int(offs[i]):int(offs[i + 1])].tobytes() for i in range(n)]


def record_names(l, path):
    buf = C.create_string_buffer(1 << 20)
    n = l.db200h_record_names(os.fsencode(path), buf, 1 << 20)
    assert 0 <= n <= (1 << 20), n
    return buf.raw[:n]


def get_paths(l, path):
    buf = C.create_string_buffer(1 << 20)
    n = l.db200h_get_paths(os.fsencode(path), buf, 1 << 20)
    assert 0 <= n <= (1 << 20), n
    return buf.raw[:n].decode().split("\n")[:-1]


def materialise_inputs(cli_npz, directory):
    names = [str(x) for x in cli_npz["names"]]
    for n in names:
        with open(os.path.join(directory, n), "wb") as f:
            f.write(cli_npz["file_" + n].tobytes())
    return names


_NUM = re.compile(rb"^[-+]?(\d+\.?\d*([eE][-+]?\d+)?|inf|nan)$")


def assert_text_matches(got: bytes, want: bytes, rtol=2e-5, what=""):
    """Same lines, same fields; numeric fields within rtol (the reference prints 6 significant digits)."""
    gl, wl = got.split(b"\n"), want.split(b"\n")
    assert len(gl) == len(wl), f"{what}: {len(gl)} lines vs {len(wl)}\n{got[:400]!r}\n{want[:400]!r}"
    for a, b in zip(gl, wl):
        fa, fb = a.split(b"\t"), b.split(b"\t")
        assert len(fa) == len(fb), f"{what}: field count differs: {a!r} vs {b!r}"
        for x, y in zip(fa, fb):
            if x == y:
                continue
            if b":" in x and b":" in y:          # nearest-neighbour tables: "<index>:<value>"
                (xi, x), (yi, y) = x.split(b":", 1), y.split(b":", 1)
                assert xi == yi, f"{what}: neighbour index {xi!r} != {yi!r} in {a!r} vs {b!r}"
            assert _NUM.match(x.strip()) and _NUM.match(y.strip()), f"{what}: {x!r} != {y!r}"
            fx, fy = float(x), float(y)
            assert abs(fx - fy) <= rtol * max(abs(fx), abs(fy), 1e-30), f"{what}: {x!r} vs {y!r}"
