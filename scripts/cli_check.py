"""Drop-in CLI check: build an assets/ tree like the reference's (scene JSON + generated OBJ),
run `cuda_pt --output out.png scenes/bunny.json --spp 4`, compare the PNG with the API render."""
import json, os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from PIL import Image
import cuda_path_tracer_b200 as pt
from tests.test_abi_and_host import BUNNY_JSON

with tempfile.TemporaryDirectory() as tmp:
    os.makedirs(os.path.join(tmp, "assets", "scenes")); os.makedirs(os.path.join(tmp, "assets", "models"))
    mesh = pt.bunny_like(4)
    pt.write_obj(os.path.join(tmp, "assets", "models", "bunny.obj"), mesh)
    js = dict(BUNNY_JSON); js["camera"] = {"vfov": 60, "resolution": [480, 270]}
    json.dump(js, open(os.path.join(tmp, "assets", "scenes", "bunny.json"), "w"))
    work = os.path.join(tmp, "assets", "scenes")  # cwd below assets/: discovery walks up
    exe = os.path.join(ROOT, "cuda_path_tracer_b200", "cuda_pt")
    out = os.path.join(tmp, "out.png")
    r = subprocess.run([exe, "--output", out, "--spp", "4", "--max-depth", "8", "--stats-json",
                        os.path.join(tmp, "stats.json"), "scenes/bunny.json"], cwd=work, capture_output=True, text=True)
    print(r.stdout, r.stderr[-500:])
    assert r.returncode == 0
    img = np.asarray(Image.open(out))
    sd = pt.bunny_scene(mesh, 480, 270)
    tr = pt.PathTracer(max_depth=8); tr.max_iterations = 4
    tr.create_buffers((480, 270), sd); tr.render(sd.camera, 4)
    api = tr.send_to_preview()
    d = np.abs(img.astype(int) - api.astype(int))
    print("cli vs api: max diff", d.max(), "frac>1:", (d > 1).mean(), "stats:", open(os.path.join(tmp, "stats.json")).read()[:300])
    assert (d > 1).mean() < 0.01
    r = subprocess.run([exe, "scenes/bunny.json"], cwd=work, capture_output=True, text=True)
    print("no --output ->", r.returncode, r.stderr.strip())
    r = subprocess.run([exe], cwd=work, capture_output=True, text=True)
    print("no file ->", r.returncode, r.stderr.strip())
print("CLI OK")
