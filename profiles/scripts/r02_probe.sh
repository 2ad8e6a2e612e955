mkdir -p gpurun_out /tmp/rep
python profiles/scripts/shard_probe.py 8 auto
python profiles/scripts/shard_probe.py 8 uniform,on,768
python profiles/scripts/shard_probe.py 8 lane,off,256
python profiles/scripts/shard_probe.py 88 uniform,on,768 2>&1 | tail -1
python profiles/scripts/shard_probe.py 1 uniform,on,768
timeout 600 ncu --section SourceCounters --section WarpStateStats --import-source on --clock-control none -k regex:catch_at_age -s 10 -c 1 -f -o /tmp/rep/probe python profiles/scripts/shard_probe.py 8 uniform,on,768 > gpurun_out/probe_ncu.log 2>&1
ncu -i /tmp/rep/probe.ncu-rep --page source --csv --print-source sass > gpurun_out/probe_src.csv 2> gpurun_out/probe_export.log
ncu -i /tmp/rep/probe.ncu-rep --page raw --csv > gpurun_out/probe_raw.csv 2>> gpurun_out/probe_export.log
ls -la gpurun_out/probe_src.csv
