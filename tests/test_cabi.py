"""The C-ABI library builds in-tree, loads without a GPU and exports every symbol the header declares."""
import ctypes as C
import os
import re

from mizuroute_b200 import build as mrbuild
from mizuroute_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mizuroute_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mr_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_are_exported():
    lib = C.CDLL(mrbuild.build())
    names = _declared()
    assert len(names) >= 16
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(capi.EXPORTS) == names


def test_options_struct_matches_header_field_order():
    src = open(os.path.join(ROOT, "include", "mizuroute_b200.h")).read()
    body = re.sub(r"/\*.*?\*/", "", src[src.index("typedef struct {"):src.index("} mr_options;")], flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.replace("typedef struct {", "").strip()
        if not decl:
            continue
        names = decl.split(None, 1)[1]
        fields += [re.sub(r"\[.*\]", "", x).strip() for x in names.split(",")]
    assert fields == [f[0] for f in capi.mr_options._fields_]


def test_create_fails_loudly_without_a_gpu_or_with_bad_options():
    import torch
    L = capi.load()
    o = capi.mr_options()
    o.dt = 3600.0; o.n_routes = 1; o.route_methods[0] = 1; o.max_batch = 1; o.device = 0
    h = C.c_void_p()
    msg = C.create_string_buffer(256)
    if not torch.cuda.is_available():
        assert L.mr_create(C.byref(o), C.byref(h), msg) == 90 and b"no CPU path" in msg.value
    o.route_methods[0] = 6                      # routing method ids are the digits 0-5 (main_route.f90:331)
    assert L.mr_create(C.byref(o), C.byref(h), msg) == 81
    o.n_routes = 2; o.route_methods[0] = 4; o.route_methods[1] = 4      # each method at most once
    assert L.mr_create(C.byref(o), C.byref(h), msg) == 81
    o.n_routes = 1
    o.route_methods[0] = 1; o.dt = 0.0
    assert L.mr_create(C.byref(o), C.byref(h), msg) == 1
