# round 2, session 4: tree drain of the L2 kernel at 1 / 2 / 4 epilogue warps per TMEM lane quarter
O=gpurun_out/r5n; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q -k "l2" > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -3 $O/pytest.log
for g in 1 2 4; do ZS_L2_EPI_GROUPS=$g timeout 300 python tools/bench_l2.py --pairs 64 > $O/l2_tree_g$g.json 2> $O/err.txt; python -c "
import json; d=json.load(open('$O/l2_tree_g$g.json')); print('tree groups $g', d['tcgen05_i8']['ms_per_call'])"; done
for g in 1 4; do ZS_L2_CHAINS=1 ZS_L2_EPI_GROUPS=$g timeout 300 python tools/bench_l2.py --pairs 64 > $O/l2_chains_g$g.json 2> $O/err.txt; python -c "
import json; d=json.load(open('$O/l2_chains_g$g.json')); print('chains groups $g', d['tcgen05_i8']['ms_per_call'])"; done
