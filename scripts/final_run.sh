#!/bin/bash
# Round-end GPU pass, most important artefacts first: GPU tests, bench lines (configs 2 and 3), ncu captures.
O=gpurun_out
mkdir -p $O
timeout 420 python -m pytest tests -m gpu -x -q > $O/f_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/f_pytest.log
tail -3 $O/f_pytest.log
timeout 200 python bench.py > $O/f_config2.json 2> $O/f_config2.err; echo "bench rc=$?"
timeout 200 python bench.py --order 15 --nel 32 --steps 3 --warmup 3 --no-cpu > $O/f_config3.json 2> $O/f_config3.err; echo "bench15 rc=$?"
python scripts/summ.py $O/f_config2.json $O/f_config3.json
ORDER=15 NEL=16 ITERS=4 timeout 200 ncu --set full --clock-control none --import-source on -k regex:axl_kernel -c 3 \
  -o $O/f_ncu_n15 -f python scripts/profile_target.py > $O/f_ncu_n15.log 2>&1; echo "ncu15 rc=$?"
ORDER=7 NEL=32 ITERS=4 timeout 200 ncu --set full --clock-control none --import-source on -k regex:'ax8_kernel|cg_update|gs_fixed' -c 6 \
  -o $O/f_ncu_n7 -f python scripts/profile_target.py > $O/f_ncu_n7.log 2>&1; echo "ncu7 rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/f_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu > $O/f_launches.log 2>&1; echo "launches rc=$?"
ls -la $O | tail -12
