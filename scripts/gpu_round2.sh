#!/bin/bash
# First GPU call(s) of round 2 (one B200; ~30 GPU-minutes in all - split at the blank-line groups if needed):
# (1) the established GPU suite, (2) everything finished after round 1's GPU budget ran out and never run on a GPU -
# tests/test_z1..z6 (train() generator, Prodigy, dense backward, native CLIP, VAE-encode prologue, rank 64 / odd widths /
# shared dscores), (3) the default bench WITH its CPU leg (first full-size step_loss_delta), then A/Bs: native CLIP, shared
# dscores; BASELINE configs 5 (--full-ft), 2 (SD1.5) and 4 (rank 32), (4) VAE-encode timing, (5) the ncu --set full
# captures of the final pair / flash kernels.
#   gpurun --timeout 2400 -- bash scripts/gpu_round2.sh
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_z5_vae_gpu.py --deselect tests/test_z4_clip_gpu.py \
    --deselect tests/test_z3_dense_gpu.py --deselect tests/test_z2_prodigy_gpu.py --deselect tests/test_z1_train_gpu.py --deselect tests/test_z6_shapes_gpu.py \
    > gpurun_out/pytest_gpu.log 2>&1; echo "pytest (established) exit $?"; tail -3 gpurun_out/pytest_gpu.log
timeout 400 python -m pytest tests/test_z4_clip_gpu.py -q -x > gpurun_out/pytest_zclip.log 2>&1; echo "pytest zclip exit $?"; tail -15 gpurun_out/pytest_zclip.log
timeout 400 python -m pytest tests/test_z5_vae_gpu.py -q -x > gpurun_out/pytest_zvae.log 2>&1; echo "pytest zvae exit $?"; tail -15 gpurun_out/pytest_zvae.log
timeout 400 python -m pytest tests/test_z3_dense_gpu.py -q -x > gpurun_out/pytest_zdense.log 2>&1; echo "pytest zdense exit $?"; tail -15 gpurun_out/pytest_zdense.log
timeout 300 python -m pytest tests/test_z2_prodigy_gpu.py tests/test_z1_train_gpu.py tests/test_z6_shapes_gpu.py -q > gpurun_out/pytest_zprodigy_ztrain.log 2>&1; echo "pytest zprodigy+ztrain exit $?"; tail -8 gpurun_out/pytest_zprodigy_ztrain.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench default exit $?"; cut -c1-300 gpurun_out/bench_default.json; grep -o '"step_loss_delta": {[^}]*}' gpurun_out/bench_default.json
B200_NATIVE_CLIP=1 timeout 400 python bench.py --steps 10 --warmup 3 --skip-cpu --skip-roofline > gpurun_out/bench_native_clip.json 2> gpurun_out/bench_native_clip.err
echo "bench native clip exit $?"; cut -c1-300 gpurun_out/bench_native_clip.json; tail -3 gpurun_out/bench_native_clip.err
timeout 500 python bench.py --full-ft --batch 1 --steps 5 --warmup 3 --skip-cpu --no-graph > gpurun_out/bench_full_ft.json 2> gpurun_out/bench_full_ft.err
echo "bench full-ft (BASELINE config 5) exit $?"; cut -c1-300 gpurun_out/bench_full_ft.json; tail -3 gpurun_out/bench_full_ft.err
B200_SHARED_DSCORES=1 timeout 400 python bench.py --steps 10 --warmup 3 --skip-cpu --skip-roofline > gpurun_out/bench_shared_dscores.json 2> gpurun_out/bench_shared_dscores.err
echo "bench shared dscores exit $?"; cut -c1-300 gpurun_out/bench_shared_dscores.json
timeout 300 python bench.py --family sd15 --res 512 --batch 4 --rank 16 --steps 10 --warmup 3 --skip-cpu --skip-roofline > gpurun_out/bench_config2_sd15.json 2> gpurun_out/bench_config2_sd15.err
echo "bench BASELINE config 2 (SD1.5 r16 512 B4) exit $?"; cut -c1-300 gpurun_out/bench_config2_sd15.json
timeout 300 python bench.py --rank 32 --steps 10 --warmup 3 --skip-cpu --skip-roofline > gpurun_out/bench_config4_r32_1gpu.json 2> gpurun_out/bench_config4_r32_1gpu.err
echo "bench BASELINE config 4 shard (SDXL r32 B2, one of 8 ranks) exit $?"; cut -c1-300 gpurun_out/bench_config4_r32_1gpu.json
timeout 300 python scripts/time_vae.py > gpurun_out/vae_timing.txt 2>&1; cat gpurun_out/vae_timing.txt
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 3 -c 1 -o gpurun_out/gemm2_lora_full \
    python scripts/one_gemm.py 2048 1280 1280 lora > gpurun_out/ncu_full.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:flash_fwd_kernel -s 2 -c 1 -o gpurun_out/flash_fwd_full \
    python scripts/one_flash.py 4096 >> gpurun_out/ncu_full.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:flash_bwd_kernel -s 2 -c 1 -o gpurun_out/flash_bwd_full \
    python scripts/one_flash.py 4096 >> gpurun_out/ncu_full.log 2>&1
grep -a "Report\|error" gpurun_out/ncu_full.log
