#!/bin/bash
# end-of-round evidence on one GPU: the GPU test suite, the default bench line and the reference arm
O=gpurun_out
( time python -m pytest tests -m gpu -x -q ) > $O/z_tests.log 2>&1; echo tests rc=$? 
( time python bench.py ) > $O/z_n1.json 2> $O/z_n1.err; echo bench rc=$?
( time python bench.py --impl reference --steps 20 --warmup 3 ) > $O/z_ref.json 2> $O/z_ref.err; echo ref rc=$?
tail -3 $O/z_tests.log
