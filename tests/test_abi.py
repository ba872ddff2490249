"""The C-ABI boundary: the built CUDA library loads (no GPU needed for dlopen) and exports every
function include/pgtg_b200.h declares; the ctypes mirror of pgtg_config has the C layout; calls
fail loudly rather than falling back when there is no device."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pgtg_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pgtg_[a-z_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    names = declared_functions()
    for n in ("pgtg_create", "pgtg_destroy", "pgtg_reset", "pgtg_step", "pgtg_step_host", "pgtg_get_buffers", "pgtg_dlpack",
              "pgtg_load_fixed_map", "pgtg_load_draws", "pgtg_get_state", "pgtg_set_state", "pgtg_stats", "pgtg_observe"):
        assert n in names


def test_cuda_library_exports_every_declared_symbol():
    import __graft_entry__ as g

    lib = C.CDLL(g.build_cuda())
    for name in declared_functions():
        assert hasattr(lib, name), f"libpgtg_b200.so does not export {name}"
    from pgtg_b200 import _lib

    assert set(_lib.EXPORTS) == set(declared_functions())
    assert _lib.load().pgtg_abi_version() == 2


def test_config_struct_layout_matches_c(tmp_path):
    """sizeof / offsetof of pgtg_config, pgtg_rule, pgtg_tile as seen by the C compiler."""
    from pgtg_b200.config import PgtgConfig, PgtgRule, PgtgTile

    src = tmp_path / "probe.c"
    fields = ["num_envs", "env_id_base", "seed", "map_w", "obstacle_cdf", "num_channels", "sum_subgoals_reward", "traffic_density",
              "profile_cdf", "drv_min_following", "num_rules", "rules", "max_episode_steps", "max_cars"]
    body = "".join(f'printf("{f} %zu\\n", offsetof(pgtg_config, {f}));' for f in fields)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "pgtg_b200.h"\nint main(){printf("config %zu\\nrule %zu\\ntile %zu\\n", sizeof(pgtg_config), sizeof(pgtg_rule), sizeof(pgtg_tile));'
                   + body + "return 0;}")
    exe = tmp_path / "probe"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    out = dict(line.split() for line in subprocess.check_output([str(exe)], text=True).splitlines())
    assert int(out["config"]) == C.sizeof(PgtgConfig) and int(out["rule"]) == C.sizeof(PgtgRule) and int(out["tile"]) == C.sizeof(PgtgTile)
    for f in fields:
        assert int(out[f]) == getattr(PgtgConfig, f).offset, f


def test_no_cpu_fallback_without_a_device():
    """Without a CUDA device the product refuses to construct an env (there is no CPU path)."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from pgtg_b200 import PGTGVectorEnv

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        PGTGVectorEnv(4)


def test_invalid_configs_are_rejected_by_the_library():
    """pgtg_create validates the POD itself (through the emulation build of the same host code)."""
    from native_env import build_emu
    from pgtg_b200 import _lib
    from pgtg_b200.config import make_config

    lib = _lib.load(build_emu())
    hc = make_config(num_envs=4)
    h = C.c_void_p()
    hc.pod.abi_version = 99
    assert lib.pgtg_create(C.byref(hc.pod), 0, C.byref(h)) == -1 and b"abi_version" in lib.pgtg_last_error()
    hc = make_config(num_envs=4)
    hc.pod.map_w = 17
    assert lib.pgtg_create(C.byref(hc.pod), 0, C.byref(h)) == -1
    hc = make_config(num_envs=4)
    assert lib.pgtg_create(C.byref(hc.pod), 0, C.byref(h)) == 0
    assert lib.pgtg_step(h, None, 4, None) == -3 and b"before reset" in lib.pgtg_last_error()  # PGTG_ERR_STATE
    lib.pgtg_destroy(h)
