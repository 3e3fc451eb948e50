"""Generates tests/golden/cameras_subset.json from the reference's shipped camera rig.

Run in the build container only (needs /root/reference):  python tests/golden/make_camera_fixture.py
Source: /root/reference/core/dataset/camera_full_calibration.json (90 views; K, R, T per view).
The subset pins sigman_release_b200.cameras.orbit_w2c (procedural) to the reference's data.
"""
import json
import os

SRC = "/root/reference/core/dataset/camera_full_calibration.json"
VIEWS = ["0000", "0008", "0029", "0030", "0037", "0045", "0053", "0059", "0060", "0065", "0082", "0085", "0089"]

if __name__ == "__main__":
    d = json.load(open(SRC))
    out = {"source": SRC, "num_views_in_source": len(d), "views": {k: {"K": d[k]["K"], "R": d[k]["R"], "T": d[k]["T"],
           "height": d[k]["height"], "width": d[k]["weight"]} for k in VIEWS}}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cameras_subset.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path)
