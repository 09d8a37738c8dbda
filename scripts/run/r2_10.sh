# round 2, GPU call 10: full -m gpu suite at HEAD (config-size parity tests, seating geometry, communicator of one)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu --durations=8 2>&1 | tail -22 | tee gpurun_out/r2_10_pytest.log
