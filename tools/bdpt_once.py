import os, sys
sys.path.insert(0, 'lightmetrica-v3_b200')
import lmcu
from lmcu import abi, scenes
mv = int(sys.argv[1]) if len(sys.argv) > 1 else 10
sc = lmcu.Scene(scenes.cornell_spheres(64))
sc.render("bdpt", abi.render_params(1920, 1080, max_verts=mv, num_samples=1920 * 1080, primary_ray_sampling_mode=abi.PRIMARY_IMAGE), download=False)
