import subprocess, sys
cases = [("float32", "(64,64)", "(1,3)"), ("float32", "(64,64)", "(1,5)"), ("float32", "(64,64)", "(3,1)"), ("int64", "(16,16)", "(1,2)"),
         ("int32", "(8,8)", "(3,1)"), ("int32", "(64,64)", "(3,3)"), ("float32", "(64,256)", "(3,3)"), ("float32", "(64,252)", "(3,7)")]
for c in cases:
    r = subprocess.run([sys.executable, "tools/dbg_tile2.py", *c], capture_output=True, text=True, env={**__import__("os").environ, "NDCONV_DEBUG_TMA": "1"})
    enc = [l for l in r.stderr.splitlines() if "tmap encode" in l]
    print(c, "PASS" if r.stdout.startswith("ok") else "FAIL", enc[-1][-60:] if enc else "no-tma")
