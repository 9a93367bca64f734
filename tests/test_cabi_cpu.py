"""The C-ABI library loads without a GPU and exports every symbol include/b200_lora.h declares; the ctypes
signatures cover exactly that set; argument errors come back as status codes + message (no exception, no crash)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "b200_lora.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    from sd_lora_trainer_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 29
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(_lib.SIGNATURES) == names, set(names) ^ set(_lib.SIGNATURES)
    assert lib.b200_version() == 1


def test_errors_are_status_codes():
    from sd_lora_trainer_b200 import _lib
    lib = _lib.load()
    assert lib.b200_gemm(None, None) != 0
    assert b"null descriptor" in lib.b200_last_error()
    d = _lib.GemmDesc()
    d.M, d.N, d.num_seg = 0, 8, 1
    assert lib.b200_gemm(ctypes.byref(d), None) != 0
    assert b"empty problem" in lib.b200_last_error()
    assert lib.b200_geglu_fwd(None, None, 4, 7, 0, None) != 0       # inner % 8 != 0
    assert lib.b200_act_fwd(None, None, 8, 0, None) != 0 and b"act_fwd" in lib.b200_last_error()          # null buffers
    assert lib.b200_act_bwd(1, 1, 1, 8, 2, None) != 0 and b"kind 2" in lib.b200_last_error()              # unknown activation
    assert lib.b200_norm_param_grad(1, 1, None, None, 1, 1, 1, 8, 4, 12, 0, 0, None) != 0                 # C % 8 != 0
    assert lib.b200_norm_param_grad(1, 1, None, None, 1, 1, 1, 10, 4, 16, 4, 0, None) != 0                # rows % hw != 0
    assert b"norm_param_grad" in lib.b200_last_error()
    assert lib.b200_norm_param_grad(1, 1, None, None, 1, 1, 1, 8, 4, 16, 0, 1, None) != 0                 # SiLU without GroupNorm
    assert lib.b200_latent_sample(None, None, None, 1.0, None, 0, None) != 0
    with pytest.raises(_lib.B200Error):
        _lib.check(2, "x")


def test_ops_refuse_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from sd_lora_trainer_b200 import _lib, ops
    with pytest.raises(_lib.B200Error):
        ops.silu_fwd(torch.zeros(8, dtype=torch.bfloat16))


def test_adamw_hyper_packing_is_host_only():
    import torch
    from sd_lora_trainer_b200 import ops
    h = torch.zeros(12)
    ops.adamw_pack_hyper(h, lr=3e-4, wd=0.004, l1_coeff=1e-9, lr2=1e-3, wd2=0.0, step=1)
    assert abs(float(h[0]) - (1 - 3e-4 * 0.004)) < 1e-7          # decay of the LoRA segment
    assert abs(float(h[1]) + 3e-4 / (1 - 0.9)) < 1e-6             # -(lr / bias_correction1) at step 1
    assert float(h[3]) == 1.0                                     # wd2 == 0 -> decay op skipped


def test_library_sass_uses_the_blackwell_paths():
    """The built sm_100a library really contains what DESIGN.md says it does (no GPU needed: cuobjdump reads the cubin):
    tcgen05.mma as CTA pairs and single CTAs, TMEM loads AND stores (the attention kernels keep P / dS in tensor memory), TMA
    loads / stores / reduce-adds, and packed fp32x2 arithmetic; mma.sync only in the batched LoRA weight-gradient kernel."""
    import shutil
    import subprocess
    import pytest
    from sd_lora_trainer_b200 import build
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", build.build()], capture_output=True, text=True, check=True).stdout
    for mnemonic in ("UTCHMMA.2CTA", "UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "FFMA2"):
        assert mnemonic in sass, mnemonic
    assert "sm_100a" in sass
    fn, users = None, set()
    for line in sass.splitlines():
        if "Function :" in line:
            fn = line.split("Function :")[1].strip()
        elif " HMMA" in line and fn:
            users.add(fn)
    assert users and all("lora_wgrad_batch" in u for u in users), users
