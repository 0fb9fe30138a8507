#!/bin/bash
timeout 900 python -m pytest tests/test_decode_mega_gpu.py -m gpu -q -k "tensor_parallel" 2>&1 | tail -3
