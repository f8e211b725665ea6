"""Quick same-box A/B: run bench.py (device-resident leg only) for each library build and scene.

    python tools/ab.py name1=path/to/lib1.so name2=path/to/lib2.so,RD3_GROUP:8   (default: the in-tree build;
    ,KEY:VALUE pairs are environment variables of that run)
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libs = [a.split("=", 1) for a in sys.argv[1:] if "=" in a] or [["tree", ""]]
scenes = [a for a in sys.argv[1:] if a in ("mixture", "ground")] or ["mixture", "ground"]
for name, spec in libs:
    for scene in scenes:
        env = dict(os.environ)
        path, *kv = spec.split(",")                      # name=lib.so,RD3_X:1,RD3_Y:2
        for a in kv:
            env[a.split(":", 1)[0]] = a.split(":", 1)[1]
        if path:
            env["RD3_LIB_PATH"] = os.path.join(ROOT, path)
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "50", "--warmup", "5",
                            "--no-cpu-baseline", "--no-e2e", "--no-rows", "--no-masks", "--scene", scene],
                           capture_output=True, text=True, env=env)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if not line:
            print(name, scene, "FAILED", r.stderr[-400:])
            continue
        j = json.loads(line[-1])
        st = j["path_roofline"]["stage_ms_per_step_single_stream"]
        print(name, scene, round(j["ms_per_step"], 3), int(j["frames_per_sec"]),
              " ".join(f"{k} {v:.3f}" for k, v in st.items()), flush=True)
