#!/bin/bash
# compute-sanitizer passes over scripts/sanitize_small.py (after a plain run of the same script exited 0) + size sweep
python scripts/sanitize_small.py > gpurun_out/san_plain.log 2>&1; echo plain rc=$?; tail -3 gpurun_out/san_plain.log
timeout 600 compute-sanitizer --tool memcheck --log-file gpurun_out/san_memcheck.txt python scripts/sanitize_small.py > gpurun_out/san_memcheck.out 2>&1; echo memcheck rc=$?; tail -5 gpurun_out/san_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/san_racecheck.txt python scripts/sanitize_small.py > gpurun_out/san_racecheck.out 2>&1; echo racecheck rc=$?; tail -5 gpurun_out/san_racecheck.txt
python scripts/sweep.py > gpurun_out/sweep_s1.json 2> gpurun_out/sweep_s1.err; echo sweep rc=$?
