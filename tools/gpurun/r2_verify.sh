# gpurun --gpus N -- 'bash tools/gpurun/r2_verify.sh N': strip tests (single-process transport) + the NCCL transport
# against one handle (lane-ordered, force-active and SourceSink crowds)
n=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_strips.py tests/test_gpu_graphs.py -x -q 2>&1 | tail -5
export NCCL_DEBUG=WARN
timeout 600 python bench.py --gpus $n --steps 60 --verify-dist > gpurun_out/r2_verify_$n.json 2> gpurun_out/r2_verify_$n.err; echo "rc=$?"
grep -v "^$\|\*\*\*\|OMP_NUM" gpurun_out/r2_verify_$n.err | tail -5; cat gpurun_out/r2_verify_$n.json
