import os, subprocess, sys, itertools
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ".")
    import raingun_b200 as rg
    from raingun_b200.synth import make_scene
    out = []
    for name in ("C3", "C4"):
        sd, spec = make_scene(name)
        sc = rg.Scene(sd); sc.set_accel(2)
        best = 1e9
        for it in range(4):
            sc.render_rows(spec.width, spec.height, 0, spec.height); best = min(best, sc.last_stats.ms_trace)
        out.append("%s trace %.2f ms" % (name, best))
        sc.close()
    print(os.environ.get("TAG"), " | ".join(out), flush=True)
else:
    combos = [dict(RG_GRID_REFILL=r, RG_GRID_QUORUM=q, RG_GRID_BURST=b, RG_GRID_BPS=p)
              for r, q, b, p in [(12,10,4,6),(8,10,4,6),(16,10,4,6),(24,10,4,6),(12,6,4,6),(12,16,4,6),(12,10,2,6),(12,10,8,6),(12,10,4,4),(12,10,4,8),(16,16,8,6),(8,6,2,6),(1,1,1,6),(32,16,8,6)]]
    for c in combos:
        env = dict(os.environ, TAG=str(c), **{k: str(v) for k, v in c.items()})
        subprocess.run([sys.executable, __file__, "child"], env=env)
