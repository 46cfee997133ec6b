cd /root/repo
mkdir -p gpurun_out
timeout 60 tools/bin/tmem_probe 2>&1 | tee gpurun_out/r02_tmem_probe.txt
bash tools/gpu_sanitize.sh
