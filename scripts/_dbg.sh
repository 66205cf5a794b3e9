python scripts/dbg_full.py 2>&1 | tail -20
echo "== memcheck"
timeout 300 compute-sanitizer --tool memcheck python scripts/dbg_full.py 2>&1 | grep -E "ERROR SUMMARY|Invalid|at 0x|by thread" | head -20
