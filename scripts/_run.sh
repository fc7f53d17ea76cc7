timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/_sanitize.py 2>&1 | tail -n 8
echo "memcheck rc=$?"
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python scripts/_sanitize.py 2>&1 | tail -n 8
