"""CPU: the C-ABI library loads and exports every symbol include/eslam_b200.h declares (no compute calls)."""
import ctypes
import os
import re

from conftest import ROOT


def header_functions():
    text = open(os.path.join(ROOT, "include", "eslam_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(eslam_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import myslam_b200._lib as L

    lib = L.load()
    names = header_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/eslam_b200.h but not exported"
    assert lib.eslam_abi_version() == L.ABI_VERSION


def test_binding_covers_header_and_arity():
    import myslam_b200._lib as L

    text = open(os.path.join(ROOT, "include", "eslam_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    for name in header_functions():
        if name in ("eslam_last_error", "eslam_abi_version", "eslam_set_debug", "eslam_exchange_flag_words",
                    "eslam_q_exchange_stage_floats", "eslam_q_touched_bytes", "eslam_mc_blocks"):
            continue
        assert name in L.PROTOTYPES, f"{name} has no ctypes prototype"
        m = re.search(name + r"\s*\((.*?)\)\s*;", text, flags=re.S)
        n_args = len([a for a in m.group(1).split(",") if a.strip()])
        assert n_args == len(L.PROTOTYPES[name]), (name, n_args, len(L.PROTOTYPES[name]))


def test_struct_layouts_match_header():
    import myslam_b200._lib as L

    assert ctypes.sizeof(L.Plane) == 16
    assert ctypes.sizeof(L.FieldDesc) == 12 * 16 + 8 + 8 + 24
    assert ctypes.sizeof(L.Camera) == 40
    assert ctypes.sizeof(L.RenderCfg) == 8 + 6 * 8
    assert L.DEC_FLOATS % 4 == 0 and L.DEC_BETA < L.DEC_FLOATS
    assert ctypes.sizeof(L.Peers) == 16 + 8 * L.MAX_PEERS + 8 + 8 + 8
    assert L.load().eslam_exchange_flag_words() == 4 * L.MAX_PEERS


def test_missing_library_fails_loudly(monkeypatch):
    import myslam_b200._lib as L

    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", "/nonexistent/libeslam_b200.so")
    try:
        L.load()
    except RuntimeError as e:
        assert "no CPU or PyTorch fallback" in str(e)
    else:
        raise AssertionError("load() must raise when the library is missing")


def test_cpu_tensors_are_rejected():
    import torch
    import myslam_b200 as M

    planes = tuple([torch.zeros(1, 32, 4, 4), torch.zeros(1, 32, 8, 8)] for _ in range(6))
    dec = M.Decoders()
    dec.bound = torch.tensor([[0., 1.], [0., 1.], [0., 1.]])
    try:
        dec(torch.zeros(5, 3), planes)
    except RuntimeError as e:
        assert "CUDA" in str(e)
    else:
        raise AssertionError("CPU tensors must be rejected: there is no CPU path")


def test_every_entry_point_rejects_null_arguments_before_touching_the_device():
    """Error behaviour of the boundary (include/eslam_b200.h: negative ESLAM_E* code + eslam_last_error, no exception,
    no device work): every entry point called with NULL pointers and zero sizes returns ESLAM_EINVAL and names itself.
    Run in a child process so that a missing check shows up as a crash of the child, not of the test session."""
    import subprocess
    import sys

    code = r"""
import ctypes as C, sys
sys.path.insert(0, %r)
import myslam_b200._lib as L
lib = L.load()
bad = []
for name, argtypes in L.PROTOTYPES.items():
    args = []
    for t in argtypes:
        if t in (C.c_int, C.c_int64):
            args.append(0)
        elif t is C.c_double:
            args.append(0.0)
        else:
            args.append(None)
    rc = getattr(lib, name)(*args)
    msg = lib.eslam_last_error().decode()
    stem = name.replace("_frames", "").replace("_act", "").replace("_sparse", "")
    if rc != -1 or not (name in msg or stem in msg or name.rsplit("_", 1)[0] in msg):
        bad.append((name, rc, msg))
print("BAD", bad)
sys.exit(1 if bad else 0)
""" % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_stream_override_nests_and_restores():
    """_lib.on_stream only redirects THIS library's launches (a module variable, no CUDA call): nested blocks restore
    the outer handle, also when an exception leaves the block."""
    from myslam_b200 import _lib

    assert _lib._STREAM_OVERRIDE is None
    with _lib.on_stream(111):
        assert _lib.stream() == 111
        with _lib.on_stream(222):
            assert _lib.stream() == 222
        assert _lib.stream() == 111
        try:
            with _lib.on_stream(333):
                raise ValueError("leave the block")
        except ValueError:
            pass
        assert _lib.stream() == 111
    assert _lib._STREAM_OVERRIDE is None
