"""Host (single-lane) build of the warp-cooperative KWT code -- test infrastructure only, see kwt_emul.cpp."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libkwt_emul.so")
_SRC = os.path.join(_HERE, "kwt_emul.cpp")
_CSRC = os.path.join(_HERE, "..", "..", "mizuroute_b200", "csrc")
_DEPS = [_SRC] + [os.path.join(_CSRC, f) for f in ("mr_kwt.cuh", "mr_kwt_scalar.cuh", "mr_dev.h", "mr_lanes.h", "mr_topo.h")]


def load_noisy(seed: int = 1):
    """The same build with pow() perturbed by at most one ulp (MR_TEST_POW_NOISE, see mr_kwt.cuh)."""
    so = os.path.join(_HERE, "libkwt_emul_noise%d.so" % seed)
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in _DEPS):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-DMR_TEST_POW_NOISE=%d" % (7919 * seed + 1),
                               "-x", "c++", "-shared", "-fPIC", "-o", so, _SRC])
    return C.CDLL(so)


def load():
    if not os.path.exists(_SO) or any(os.path.getmtime(f) > os.path.getmtime(_SO) for f in _DEPS):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-x", "c++", "-shared", "-fPIC",
                               "-o", _SO, _SRC])
    return C.CDLL(_SO)


_EULER_SO = os.path.join(_HERE, "libeuler_emul.so")
_EULER_SRC = os.path.join(_HERE, "euler_emul.cpp")
_EULER_DEPS = [_EULER_SRC, os.path.join(_HERE, "da_emul.h")] + [os.path.join(_CSRC, f) for f in ("mr_euler.cuh", "mr_dev.h", "mr_lanes.h", "mr_topo.h")]


def load_euler(noise_seed: int = 0):
    """Host build of the Euler routing schemes (mr_euler.cuh), see euler_emul.cpp; noise_seed > 0: pow() perturbed by at
    most one ulp (MR_TEST_POW_NOISE, mr_dev.h)."""
    so = _EULER_SO if not noise_seed else os.path.join(_HERE, "libeuler_emul_noise%d.so" % noise_seed)
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in _EULER_DEPS):
        extra = ["-DMR_TEST_POW_NOISE=%d" % (7919 * noise_seed + 1)] if noise_seed else []
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-Wno-unknown-pragmas", *extra, "-x", "c++", "-shared",
                               "-fPIC", "-o", so, _EULER_SRC])
    return C.CDLL(so)


_LAKE_SO = os.path.join(_HERE, "liblake_emul.so")
_LAKE_SRC = os.path.join(_HERE, "lake_emul.cpp")
_LAKE_DEPS = [_LAKE_SRC] + [os.path.join(_CSRC, f) for f in ("mr_lake.cuh", "mr_euler.cuh", "mr_dev.h", "mr_lanes.h", "mr_topo.h", "mr_calendar.h", "mr_lakeparams.h")]


def load_lake():
    """Host build of the lake reach step (mr_lake.cuh) inside a kinematic-wave network, see lake_emul.cpp."""
    if not os.path.exists(_LAKE_SO) or any(os.path.getmtime(f) > os.path.getmtime(_LAKE_SO) for f in _LAKE_DEPS):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-Wno-unknown-pragmas", "-x", "c++", "-shared", "-fPIC",
                               "-o", _LAKE_SO, _LAKE_SRC])
    return C.CDLL(_LAKE_SO)


_IRF_SO = os.path.join(_HERE, "libirf_emul.so")
_IRF_SRC = os.path.join(_HERE, "irf_emul.cpp")
_IRF_DEPS = [_IRF_SRC, os.path.join(_HERE, "da_emul.h")] + [os.path.join(_CSRC, f) for f in ("mr_irf.cuh", "mr_lake.cuh", "mr_dev.h", "mr_lanes.h", "mr_topo.h", "mr_uh.h")]


def load_irf():
    """Host build of the accumulation / IRF reach steps (mr_irf.cuh), see irf_emul.cpp."""
    if not os.path.exists(_IRF_SO) or any(os.path.getmtime(f) > os.path.getmtime(_IRF_SO) for f in _IRF_DEPS):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-Wno-unknown-pragmas", "-x", "c++", "-shared", "-fPIC",
                               "-o", _IRF_SO, _IRF_SRC])
    return C.CDLL(_IRF_SO)


def load_calendar():
    """Host build of mr_calendar.h, see calendar_emul.cpp."""
    so, src = os.path.join(_HERE, "libcalendar_emul.so"), os.path.join(_HERE, "calendar_emul.cpp")
    deps = [src, os.path.join(_CSRC, "mr_calendar.h")]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-x", "c++", "-shared", "-fPIC", "-o", so, src])
    return C.CDLL(so)


def load_ingest():
    """Host build of the forcing ingest (mr_ingest.h), see ingest_emul.cpp."""
    so, src = os.path.join(_HERE, "libingest_emul.so"), os.path.join(_HERE, "ingest_emul.cpp")
    deps = [src, os.path.join(_CSRC, "mr_ingest.h"), os.path.join(_CSRC, "mr_dev.h")]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-x", "c++", "-shared", "-fPIC", "-o", so, src])
    return C.CDLL(so)


def load_fastpow():
    """Host build of mr_pow04_fast (mr_dev.h), see fastpow_emul.cpp."""
    so, src = os.path.join(_HERE, "libfastpow_emul.so"), os.path.join(_HERE, "fastpow_emul.cpp")
    deps = [src, os.path.join(_CSRC, "mr_dev.h"), os.path.join(_CSRC, "mr_lanes.h")]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-x", "c++", "-shared", "-fPIC", "-o", so, src])
    return C.CDLL(so)
