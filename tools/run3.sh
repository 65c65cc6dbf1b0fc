exec > gpurun_out/sanitizer_r1.log 2>&1
compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fast_float_widths or edge_cases or long_rows or tensor_core" 2>&1 | tail -12
echo "exit=$?"
