"""Static checks on the SASS of the built library (no GPU needed: cuobjdump reads the cubin).

The FP64 cost model in DESIGN.md section 5 rests on two facts about the unrolled control interval of the
benchmarked kernel, cl::k_rollout_sm<cl::EnvLorenzRK4<double>>: it issues 45 FP64-pipe instructions
per RK4 substep (24 RHS + 9 stage + 12 combine), and more than half of its 128 three-register
DFMAs find one source in the operand reuse cache.  Both depend on ptxas scheduling choices that a
change of launch bounds silently undoes (kernels_common.cuh, CL_SM_THREADS), hence this test."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

KERNELS = {
    "f64": "_ZN2cl12k_rollout_smINS_12EnvLorenzRK4IdEEEEvNS_7KParamsE14CUtensorMap_st",
    "f32": "_ZN2cl12k_rollout_smINS_12EnvLorenzRK4IfEEEEvNS_7KParamsE14CUtensorMap_st",
}


def _sass(chaos_lib, fun):
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    from gym_lorenz_b200 import build
    r = subprocess.run([exe, "-sass", "-fun", fun, build.LIB], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return r.stdout.splitlines()


def test_unrolled_interval_of_the_benchmarked_kernel(chaos_lib):
    import sass_mix
    blocks = sorted(sass_mix.blocks(_sass(chaos_lib, KERNELS["f64"])), key=len, reverse=True)
    c, three, hits = sass_mix.mix(blocks[0])
    assert c["DFMA"] == 16 * 37, c["DFMA"]                 # 16 substeps x (16 RHS + 9 stage + 12 combine)
    assert 16 * 8 <= c["DADD"] <= 16 * 8 + 12, c["DADD"]   # 2 per RHS evaluation + the interval's epilogue
    assert three == 16 * 8                                 # dy, dz of every evaluation: inherently 3 registers
    assert hits >= 96, f"only {hits} of {three} three-register DFMAs hit the operand reuse cache"
    fp64 = c["DFMA"] + c["DADD"] + c["DMUL"] + c["DSETP"]
    assert len(blocks[0]) - fp64 - c["F2F"] <= 90          # non-FP64 instructions of the interval's main block


def test_uses_bulk_async_copies_and_mbarriers(chaos_lib):
    txt = "\n".join(_sass(chaos_lib, KERNELS["f64"]))
    assert "UTMALDG.3D" in txt      # a chunk's actions: one 3-D tensor copy
    assert "UBLKCP" in txt and "SYNCS" in txt
    assert "MEMBAR" not in txt and "CCTL.IVALL" not in txt.replace("SYNCS.CCTL.IVALL", "")   # no fence in the task path


def test_f32_kernel_substep_is_45_fma_pipe_instructions(chaos_lib):
    import sass_mix
    blocks = sorted(sass_mix.blocks(_sass(chaos_lib, KERNELS["f32"])), key=len, reverse=True)
    c, _, _ = sass_mix.mix(blocks[0])
    # ptxas turns part of the additions into FFMAs; what counts is the total on the FMA pipe
    assert 16 * 45 <= c["FFMA"] + c["FADD"] + c["FMUL"] <= 16 * 45 + 14, (c["FFMA"], c["FADD"], c["FMUL"])
