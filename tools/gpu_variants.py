import os, subprocess, sys
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ".")
    import raingun_b200 as rg
    from raingun_b200.synth import make_scene
    sd, spec = make_scene("C4")
    w, h = 1920, 1080
    sc = rg.Scene(sd); sc.set_accel(int(os.environ.get("RG_ACCEL", "1")))
    best = None
    for it in range(3):
        sc.render_image(w, h); st = sc.last_stats
        if best is None or st.ms_trace < best.ms_trace: best = st
    print("variant", os.environ.get("RG_BRUTE_VARIANT", "0"), "trace %.2f ms dev %.2f ms" % (best.ms_trace, best.ms_device),
          "tests/s %.3g" % (best.body_tests / (best.ms_trace * 1e-3)), "rays", best.rays, flush=True)
else:
    for v in sys.argv[1:] or ["0", "1", "2", "3", "4", "5", "6", "7"]:
        env = dict(os.environ, RG_BRUTE_VARIANT=v)
        subprocess.run([sys.executable, __file__, "child"], env=env)
