cd /root/repo
python -m pytest tests/test_gpu_classic.py tests/test_gpu_virtual.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -3
python bench.py --skip_extra > gpurun_out/r2_b37.json 2>/dev/null
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_b37.json'))
print('ms/step',d['ms_per_step'],'eager',d['extra']['eager']['ms_per_step'],'residue',d['extra'].get('residue_pipeline',{}).get('ms_per_step'), d['extra'].get('residue_pipeline',{}).get('stages_ms'))
P
python profiles/r02_kernels.py fc 20000 | head -1
