#!/usr/bin/env python
"""Convert the Fermat scenes named by BASELINE.json into binary snapshots (.fbs) under scenes/_cache/.

The GPU boxes have no /root/reference, so the scenes travel as pre-processed snapshots of exactly the
arrays the renderer consumes (our own format, fermat_b200/csrc/host/scene.cpp save_scene_snapshot):
unified vertices with packed normals, fp16 texcoord triangles, MeshMaterial table, float4 textures,
camera. scenes/_cache/ is git-ignored (ships with gpurun); the small CornellBox snapshot is also
committed under tests/golden/ so that the GPU tests never depend on the cache.
"""
import os
import shutil
import sys
import zipfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("FERMAT_REFERENCE", "/root/reference")
CACHE = os.path.join(ROOT, "scenes", "_cache")


DIRLIGHT_FA = """# CornellBox-JP lit by its area light and two directional lights (test fixture for pathtracer_core.h:870-988)
Camera persp eye 0 1.3 1.5 aim -0.01 0.945 -0.025 up 0 1 0 fov 1.81
LoadScene %s
DirectionalLight direction 0.3 -0.4 -1.0 color 2.0 1.9 1.6
DirectionalLight dir -0.5 -0.3 -1.0 color 0.4 0.5 0.9
"""


def snapshot(args, out):
    import fermat_b200 as fb
    if not os.path.exists(out):
        sc = fb.Scene(args + ["-r", "64", "64"])      # resolution only sizes the VPL set, which is not stored
        sc.save_snapshot(out)
        print("wrote", out, sc.bvh_stats())
        sc.close()
    # the GPU boxes receive the xz twin (.gpurunignore drops the raw file); fermat_b200.resolve_scene unpacks it
    if out.startswith(CACHE) and not os.path.exists(out + ".xz"):
        import lzma
        with open(out, "rb") as f, lzma.open(out + ".xz", "wb", preset=1) as g:
            g.write(f.read())
    return out


def main(which=None):
    os.makedirs(CACHE, exist_ok=True)
    models = os.path.join(REF, "models")
    done = {}
    if not os.path.isdir(models):
        print("reference models not found at", models, "- nothing to do")
        return done
    # C1: CornellBox (camera-frontal.txt)
    cb = os.path.join(models, "CornellBox")
    done["cornellbox"] = snapshot(["-i", os.path.join(cb, "CornellBox-JP.obj"), "-c", os.path.join(cb, "camera-frontal.txt")],
                                  os.path.join(ROOT, "tests", "golden", "cornellbox_jp.fbs"))
    done["cornellbox_glossy"] = snapshot(["-i", os.path.join(cb, "CornellBox-Glossy.obj"), "-c", os.path.join(cb, "camera-frontal.txt")],
                                         os.path.join(CACHE, "cornellbox_glossy.fbs"))
    # a12 fixture: CornellBox with two DirectionalLights shining in through the open front (grammar: src/mesh/fermat_loader.cpp:294-349;
    # the reference's own example is models/bathroom2/bathroom_cornell.fa). Written next to the .obj's directory on the search path.
    fa = os.path.join(CACHE, "_cornellbox_dirlight.fa")
    with open(fa, "w") as f:
        f.write(DIRLIGHT_FA % os.path.join(cb, "CornellBox-JP.obj"))
    done["cornellbox_dirlight"] = snapshot(["-i", fa], os.path.join(ROOT, "tests", "golden", "cornellbox_dirlight.fbs"))
    os.remove(fa)
    if which == "small":
        return done
    # C4: water_caustic
    wc = os.path.join(models, "water_caustic")
    done["water_caustic"] = snapshot(["-i", os.path.join(wc, "water_caustic.fa")], os.path.join(CACHE, "water_caustic.fbs"))
    # C3: material-testball (pbrt scene: PLY meshes, substrate/glass/metal/matte materials, env-map light)
    mt = os.path.join(models, "material-testball")
    done["material_testball"] = snapshot(["-i", os.path.join(mt, "scene.pbrt")], os.path.join(CACHE, "material_testball.fbs"))
    # C2 / C5: bathroom2 (the .obj ships zipped)
    out = os.path.join(CACHE, "bathroom2.fbs")
    if not os.path.exists(out):
        work = os.path.join(CACHE, "_bathroom2_src")
        src = os.path.join(models, "bathroom2")
        if not os.path.exists(os.path.join(work, "bathroom.obj")):
            os.makedirs(work, exist_ok=True)
            with zipfile.ZipFile(os.path.join(src, "bathroom.zip")) as z:
                z.extractall(work)
            for f in ("bathroom.fa", "bathroom.mtl"):
                shutil.copy(os.path.join(src, f), os.path.join(work, f))
            if not os.path.exists(os.path.join(work, "textures")):
                shutil.copytree(os.path.join(src, "textures"), os.path.join(work, "textures"))
        done["bathroom2"] = snapshot(["-i", os.path.join(work, "bathroom.fa")], out)
        shutil.rmtree(work, ignore_errors=True)
    else:
        done["bathroom2"] = out
    return done


if __name__ == "__main__":
    main(*sys.argv[1:])
