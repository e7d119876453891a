"""CPU: the C-ABI library loads, exports every symbol include/sbq.h declares, and refuses to compute
without a device (no CPU fallback in the product)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols(header="sbq.h"):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sbq_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_all_exported(sbq_lib_path):
    from strawberry_b200 import api
    L = ctypes.CDLL(sbq_lib_path)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/sbq.h but not exported by libsbq.so"
    assert sorted(api.ABI_SYMBOLS) == declared
    from strawberry_b200 import builder
    declared_b = _declared_symbols("sbq_builder.h")
    for name in declared_b:
        assert hasattr(L, name), f"{name} declared in include/sbq_builder.h but not exported by libsbq.so"
    assert sorted(builder.BUILDER_SYMBOLS) == declared_b


def test_defaults_follow_the_reference(sbq_lib_path):
    from strawberry_b200 import api
    cfg = api.default_config()
    assert cfg.max_iter == 1000          # include/estimate.hpp:236
    assert cfg.theta_tol == 1e-2         # include/estimate.hpp:240
    assert cfg.row_eps == 1e-5           # src/estimate.cpp:381
    assert cfg.bias_mode == 0 and cfg.effective_len_norm == 0
    assert api.lib().sbq_abi_version() == 2 and cfg.n_gpus == 1
    assert b"no CUDA device" in api.lib().sbq_error_string(api.SBQ_ERR_NO_DEVICE)


def test_no_device_means_error_not_fallback(sbq_lib_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    from strawberry_b200 import api
    with pytest.raises(api.SbqError) as e:
        api.Quantifier()
    assert e.value.code == api.SBQ_ERR_NO_DEVICE


def test_product_does_not_touch_the_oracle():
    """The product package must never import / link oracle/."""
    pkg = os.path.join(ROOT, "strawberry_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text and "libsbref" not in text and "sbq_oracle" not in text, f


def test_shipped_cubin_is_sm100a_and_uses_the_blackwell_features(sbq_lib_path):
    """Static check of the shipped library (no GPU needed): the only device code is an sm_100a cubin, no kernel spills to
    local memory, the giant-locus kernels stage their rows with 1-D TMA bulk copies (SASS UBLKCP) behind mbarriers (SYNCS),
    and the cluster-tier kernels (EM and bias) use the hardware cluster barrier (UCGABAR_*). tools/sass_summary.py prints
    the full table (profiles/r02_sass_summary.txt)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    elf = subprocess.run([cuobjdump, "-lelf", sbq_lib_path], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", elf))
    assert archs == {"sm_100a"}, archs
    res = subprocess.run([cuobjdump, "-res-usage", sbq_lib_path], capture_output=True, text=True).stdout
    local = [int(x) for x in re.findall(r"LOCAL:(\d+)", res)]
    assert local and max(local) == 0, "a kernel spills to local memory"
    sass = subprocess.run([cuobjdump, "-sass", sbq_lib_path], capture_output=True, text=True).stdout
    per_kernel, cur = {}, None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per_kernel[cur] = set()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if cur and m:
            per_kernel[cur].add(m.group(1))
    def ops(substr):
        return [v for k, v in per_kernel.items() if substr in k]
    assert ops("em_grid_dual_kernel") and all({"UBLKCP", "SYNCS"} <= o for o in ops("em_grid_dual_kernel"))
    assert ops("em_grid_tma_kernel") and all({"UBLKCP", "SYNCS"} <= o for o in ops("em_grid_tma_kernel"))
    assert ops("em_cluster_kernel") and all({"UCGABAR_ARV", "UCGABAR_WAIT"} <= o for o in ops("em_cluster_kernel"))
    assert ops("em_bias_kernel") and all({"UCGABAR_ARV", "UCGABAR_WAIT"} <= o for o in ops("em_bias_kernel"))
    assert not any("HMMA" in o or "IMMA" in o for o in per_kernel.values())    # sparse gather + reduction: no tensor-core path by design
