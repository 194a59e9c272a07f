#!/bin/bash
# SURVEY.md 7.1 step 0 / VERDICT r01 item 1: can the REAL environment stack
# (pycolab -> ai-safety-gridworlds -> safe-grid-gym, reference setup.py:46,
# train.py:51-52) be obtained on the GPU box?  Records every attempt; exits 0
# either way.  Run:  gpurun -- 'bash scripts/env_probe.sh > gpurun_out/r02_env_probe.log 2>&1'
echo "== date / host"; date -u; uname -a
echo "== python"; python -V; which python
echo "== import probe"
for m in pycolab ai_safety_gridworlds safe_grid_gym gym gymnasium tensorboardX ray absl; do
    python - <<EOF
import importlib
try:
    mod = importlib.import_module("$m")
    print("$m: PRESENT", getattr(mod, "__version__", ""), getattr(mod, "__file__", ""))
except Exception as exc:
    print("$m: ABSENT (%s: %s)" % (type(exc).__name__, exc))
EOF
done
echo "== filesystem search (site-packages, /opt, /usr, /root, /home, /tmp, /workspace)"
find / -xdev \( -iname '*pycolab*' -o -iname '*ai_safety_gridworlds*' -o -iname '*ai-safety-gridworlds*' \
    -o -iname '*safe_grid_gym*' -o -iname '*safe-grid-gym*' -o -iname 'boat_race*' -o -iname 'safety_game*' \
    -o -iname 'side_effects_sokoban*' -o -iname 'tomato_watering*' \) \
    -not -path '*/proc/*' -not -path "${GRAFT_REPO_ROOT:-/nonexistent}/*" -not -path '/root/repo/*' 2>/dev/null | head -50
echo "(end of search)"
echo "== wheelhouse"
ls /opt/wheelhouse 2>/dev/null | grep -i -E 'pycolab|safety|grid|gym' || echo "no matching wheel in /opt/wheelhouse"
echo "== pip download (expects: no network)"
timeout 60 python -m pip download --no-deps -d /tmp/probe_dl pycolab 2>&1 | tail -4
echo "== pip install from the offline wheelhouse"
timeout 60 python -m pip install --no-index --find-links /opt/wheelhouse --target /tmp/probe_t pycolab 2>&1 | tail -3
timeout 60 python -m pip install --no-index --find-links /opt/wheelhouse --target /tmp/probe_t gym 2>&1 | tail -3
echo "== git clone (expects: no network)"
timeout 30 git clone --depth 1 https://github.com/deepmind/pycolab /tmp/probe_pycolab 2>&1 | tail -2
timeout 30 git clone --depth 1 https://github.com/david-lindner/safe-grid-gym /tmp/probe_sgg 2>&1 | tail -2
echo "== verdict"
python - <<'EOF'
import importlib
ok = []
for m in ("pycolab", "ai_safety_gridworlds", "safe_grid_gym"):
    try:
        importlib.import_module(m); ok.append(m)
    except Exception:
        pass
print("REAL ENV STACK AVAILABLE:" if len(ok) == 3 else "REAL ENV STACK NOT AVAILABLE; importable:", ok)
EOF
exit 0
