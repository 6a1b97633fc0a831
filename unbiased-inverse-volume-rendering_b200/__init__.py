"""B200-native differential volumetric path tracer behind the reference's IntegratorConfig /
integrator-plugin surface (hot path of rgl-epfl/unbiased-inverse-volume-rendering)."""
from . import _native, losses
from ._native import NativeError, tea32
from .integrator import (INTEGRATORS, NeRFIntegrator, Scene, VolpathSimpleIntegrator, load_dict,
                         register_integrator, render)
from .opt_config import IntegratorConfig, OptimizationConfig, Schedule, add_int_config, get_int_config
from .scene_config import SceneConfig, add_scene_config, add_scene_config_variant, get_scene_config
from .exr import read_exr, write_exr
from .rgbe import read_hdr, write_hdr
from .batched import gather_ref_values, render_batch, sample_batch_pixels, sensor_table
from .multires import (adjust_majorant_res_factor, read_vol, save_params, upsample_grid, upsample_iterations,
                       upsample_params, write_vol)
from .fd import fd_gradients
from .optimize import (PCG32, SGD, Adam, checkpoint_prefix, create_checkpoint, get_reference_image_paths,
                       initial_resolution, initialize_scene, l1_loss_grad, learning_rates, load_reference_images,
                       make_view_lanes, optimization_step, param_bounds, preview_suffix, reference_pass_plan, render_previews,
                       render_reference_image, run_optimization)
from .scene import (EnvMap, Sensor, VolumeScene, benchmark_scene, circle_sensors, cube_test_grids,
                    cube_test_scene, look_at, synthetic_grids)

__all__ = [
    "NativeError", "tea32", "losses", "INTEGRATORS", "NeRFIntegrator", "Scene", "VolpathSimpleIntegrator", "load_dict",
    "register_integrator", "render", "IntegratorConfig", "add_int_config", "get_int_config",
    "Sensor", "VolumeScene", "benchmark_scene", "circle_sensors", "cube_test_grids",
    "cube_test_scene", "look_at", "synthetic_grids",
    "gather_ref_values", "render_batch", "sample_batch_pixels", "sensor_table",
    "adjust_majorant_res_factor", "read_vol", "save_params", "upsample_grid", "upsample_iterations",
    "upsample_params", "write_vol",
    "OptimizationConfig", "Schedule", "SceneConfig", "add_scene_config", "add_scene_config_variant", "get_scene_config",
    "read_exr", "write_exr", "read_hdr", "write_hdr", "EnvMap", "PCG32", "SGD", "checkpoint_prefix", "create_checkpoint", "get_reference_image_paths",
    "initial_resolution", "initialize_scene", "load_reference_images", "preview_suffix", "reference_pass_plan",
    "render_previews", "render_reference_image", "run_optimization",
    "fd_gradients", "Adam", "l1_loss_grad", "learning_rates", "make_view_lanes", "optimization_step", "param_bounds",
]
