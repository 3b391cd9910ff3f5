cd /root/repo; mkdir -p gpurun_out
timeout 120 python tools/graph_timeline.py > gpurun_out/graph_timeline_run.log 2>&1; echo "timeline exit=$?"; ls gpurun_out | grep -i timeline
cd tools; timeout 300 ncu --set full --clock-control none --import-source on -k regex:upscore8 -c 1 -f -o ../gpurun_out/r02_upscore python fcn_timeline.py > ../gpurun_out/ncu_upscore.log 2>&1; echo "ncu exit=$?"
