# final check of the tree: bench.py with its defaults (own arm), smoke()
out=gpurun_out/${1:-r02z}; mkdir -p $out
timeout 600 python bench.py > $out/bench.json 2> $out/err.log; cut -c1-420 $out/bench.json; tail -n 3 $out/err.log
(timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1) | tee $out/smoke.log
