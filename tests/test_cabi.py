"""CPU tests of the boundary: the C-ABI library builds for sm_100a, loads, and exports
every symbol include/ic3net_b200.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "ic3net_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ic3_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    from ic3net_b200 import _lib
    assert header_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    for name in header_symbols():
        assert hasattr(lib, name), name
    lib.ic3_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.ic3_version()
    lib.ic3_strerror.restype = ctypes.c_char_p
    assert b"NULL" in lib.ic3_strerror(-1)


def test_struct_sizes_match_header(built_lib):
    """ctypes mirrors must have the C layout (checked with a tiny gcc probe)."""
    import subprocess
    import tempfile
    from ic3net_b200 import _lib
    names = {"ic3_pp_cfg": _lib.PPCfg, "ic3_pp_state": _lib.PPState, "ic3_rollout_io": _lib.RolloutIO,
             "ic3_tj_cfg": _lib.TJCfg, "ic3_tj_state": _lib.TJState, "ic3_policy_cfg": _lib.PolicyCfg,
             "ic3_policy_params": _lib.PolicyParams, "ic3_policy_packed": _lib.PolicyPacked,
             "ic3_policy_io": _lib.PolicyIO, "ic3_bptt_plan": _lib.BpttPlan, "ic3_bptt_step_io": _lib.BpttStepIO}
    prog = '#include <stdio.h>\n#include "ic3net_b200.h"\nint main(){' + "".join(
        'printf("%s %%zu\\n", sizeof(%s));' % (n, n) for n in names) + "return 0;}"
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "p.c")
        open(c, "w").write(prog)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", os.path.join(d, "p")])
        out = subprocess.check_output([os.path.join(d, "p")]).decode().split("\n")
    sizes = dict(l.split() for l in out if l)
    for n, cls in names.items():
        assert int(sizes[n]) == ctypes.sizeof(cls), n


def test_no_cpu_fallback():
    """Product modules must refuse to run without CUDA instead of falling back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import argparse
    from ic3net_b200.predator_prey_env import PredatorPreyEnv
    a = argparse.Namespace(dim=5, vision=0, moving_prey=False, mode="mixed", enemy_comm=False, nenemies=1,
                           nfriendly=3, nagents=3, no_stay=False, nenvs=2, seed=0)
    with pytest.raises(RuntimeError):
        PredatorPreyEnv().multi_agent_init(a)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "ic3net_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f
