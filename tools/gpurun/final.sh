# end-of-round check: GPU suite, smoke, default bench line, reference arm
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time timeout 300 python bench.py ) > gpurun_out/final_default.json 2> gpurun_out/final_default.err; tail -4 gpurun_out/final_default.err | head -2
python -c "
import json; d=json.loads(open('gpurun_out/final_default.json').readlines()[-1]); print('value %.4e  ms %.4f  kernel %.4f  e2e %.4e  cpu %.4e  frac %.4f  launches %d  clocks %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'], d['cpu_baseline']['value'], d['roofline']['frac'], d['gpu_launches'], d['clocks']))"
( time timeout 300 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; tail -4 gpurun_out/final_ref.err | head -2
python -c "
import json; d=json.loads(open('gpurun_out/final_ref.json').readlines()[-1]); print('reference arm value %.4e  ms/step %.2f' % (d['value'], d['ms_per_step']))"
