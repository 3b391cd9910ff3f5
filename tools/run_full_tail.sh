cd /root/repo
bash tools/r2_full.sh
timeout 300 ncu --set full --clock-control none --import-source on -k regex:classmap_tail -c 1 -f -o gpurun_out/r02_tail python tools/deconv_timeline.py > gpurun_out/ncu_tail.log 2>&1; echo "ncu tail exit=$?"
