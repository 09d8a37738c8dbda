# round 2, GPU call 50: after the fill + rejection fix in the windowed playout (found by tests/test_windowed_playout_logic.py on the CPU)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu -k "rollout or lanes or mapping or config_sizes or fpu" 2>&1 | tail -2 | tee gpurun_out/r2_50_tests.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 600 python bench.py --config 0 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('config0', d['value'], d['e2e']['value'])"
