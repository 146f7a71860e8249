#!/bin/bash
# compute-sanitizer memcheck over small invocations of every kernel family (one tool per gpurun call).
mkdir -p gpurun_out
run() {
  echo "== $*"
  "$@" > /dev/null 2>&1 || { echo "plain run failed: $*"; return; }
  timeout 900 compute-sanitizer --tool memcheck --target-processes all "$@" 2>&1 | grep -E "ERROR SUMMARY|Invalid|out of bounds|misaligned|smoke ok|us/step|chain-sweeps" | head -8
}
run python -c "import __graft_entry__ as g; g.smoke()"
run python profiles/em_only.py 20000 1 10
run python profiles/em_only.py 6000 1 20
run python profiles/rj_only.py 4096 10 c5_rj
run python profiles/rj_only.py 4096 10 toy2
