ncu --set full --clock-control none --import-source on -k regex:k_search -s 2 -c 1 -o gpurun_out/prof_search_r1h python bench.py --steps 2 --warmup 1 --no-locate --no-cpu-baseline > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_|Device" -c 800 --csv --log-file gpurun_out/launches_r1h.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/*r1h*
