timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/rec_tests.log 2>&1; tail -4 gpurun_out/rec_tests.log
python profiles/scripts/feat_time.py c5 2>&1 | grep -v Warn; ISOKANN_FEAT_BULK=0 python profiles/scripts/feat_time.py c5 2>&1 | grep -v Warn; python profiles/scripts/feat_time.py c4 2>&1 | grep -v Warn
timeout 120 python bench.py --profile --steps 1 --warmup 0 --N 131072 > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.err && timeout 300 ncu --set full --clock-control none --import-source on -k regex:featurize_blk -s 2 -c 1 -f -o gpurun_out/frec_v7 python bench.py --profile --steps 1 --warmup 0 --N 131072 > gpurun_out/ncu_frec.log 2>&1
grep -c PROF gpurun_out/ncu_frec.log
