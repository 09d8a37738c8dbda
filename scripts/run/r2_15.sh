mkdir -p gpurun_out
timeout 1200 python -X faulthandler -m pytest tests -x -q -m gpu 2>&1 | grep -v "^  File \"/opt" | tail -40 | tee gpurun_out/r2_15_pytest.log
