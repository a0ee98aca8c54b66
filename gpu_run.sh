cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_cli.py -x -q -m gpu 2>&1 | tail -3
python - <<'PY'
import sys,time,subprocess,os
sys.path.insert(0,'.')
from vcfx_b200 import synth
data=synth.make_vcf(2,200000,2504,seed=2)
open('/dev/shm/t.vcf','wb').write(data); print(len(data)/1e9,"GB")
env=dict(os.environ,VCFX_TIMING="1")
for tool,args in (("allele_freq_calc",["-q","-i"]),("hwe_tester",["-q","-i"]),("variant_counter",[]),("missing_detector",["-q","-i"])):
    for rep in range(3):
        t=time.perf_counter(); r=subprocess.run([f"vcfx_b200/bin/VCFX_{tool}",*args,"/dev/shm/t.vcf"],stdout=subprocess.DEVNULL,stderr=subprocess.PIPE,env=env); dt=time.perf_counter()-t
        print(tool,"%.2f s"%dt, r.stderr.decode().strip()[-200:])
os.unlink('/dev/shm/t.vcf')
PY
