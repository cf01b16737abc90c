"""Convert the reference's velocity table (input DATA, not source) into a compact binary fixture.

Run in the authoring container only (needs /root/reference):
    python tools/make_ak135_fixture.py
Writes raytracer.jl_b200/data/ak135_profile.npz with float64 arrays depth_km, vp, vs read from
/root/reference/VelocityProfiles/R_Vp_Vs_AK135.txt (the IASP91 file of the reference is byte-identical).
"""
import hashlib
import os
import numpy as np

SRC = "/root/reference/VelocityProfiles/R_Vp_Vs_AK135.txt"
DST = os.path.join(os.path.dirname(__file__), "..", "raytracer.jl_b200", "data", "ak135_profile.npz")

if __name__ == "__main__":
    raw = open(SRC, "rb").read()
    tab = np.loadtxt(SRC, dtype=np.float64)
    assert tab.shape == (6372, 3)
    np.savez_compressed(DST, depth_km=tab[:, 0].copy(), vp=tab[:, 1].copy(), vs=tab[:, 2].copy(),
                        source_md5=np.array(hashlib.md5(raw).hexdigest()))
    print("wrote", os.path.normpath(DST), tab.shape)
