"""
ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Loads the UNMODIFIED reference (plai-group/latent-flexible-video-diffusion-modeling) next to this repo's drop-in, so that
tests and the bench's reference / gpu_eager arms can run BOTH in one process:

  * `find_ref()`            — where the reference lives: oracle/_ref (the copy `oracle/make_ref.sh` ships to the GPU box),
                              else $FDM_REFERENCE or /root/reference (build container only); None if neither exists.
  * `load_reference()`      — the reference's `improved_diffusion` package imported under the ALIAS `fdm_ref_improved_diffusion`
                              (its hot-path modules use relative imports only, so they resolve inside the alias and never
                              collide with this repo's `improved_diffusion`).
  * `enable_mixed_package()`— the drop-in deployment: hot-path modules from this repo, everything else (train_util,
                              sampling_schemes, dist_util, logger, ...) from the reference, through the product's own
                              FDM_REFERENCE_PATH mechanism (improved_diffusion/__init__.py), plus the mpi4py / blobfile /
                              imageio stand-ins of oracle/stubs (absent from the image).
  * `load_script(name)`     — a reference script (scripts/video_sample.py, ...) as a module (its __main__ block is guarded).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / gpu_eager / --impl reference legs import this.
"""
import importlib
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ALIAS = "fdm_ref_improved_diffusion"


def find_ref():
    for cand in (os.path.join(HERE, "_ref"), os.environ.get("FDM_REFERENCE") or "", "/root/reference"):
        if cand and os.path.isfile(os.path.join(cand, "improved_diffusion", "unet.py")):
            return cand
    return None


def require_ref():
    ref = find_ref()
    if ref is None:
        raise RuntimeError("the reference is not available: run oracle/make_ref.sh (or __graft_entry__.build()) in the build "
                           "container so that oracle/_ref/ travels to the GPU box")
    return ref


def add_stubs():
    """mpi4py / blobfile / imageio stand-ins (only when the real package is absent)."""
    stubs = os.path.join(HERE, "stubs")
    for name in ("mpi4py", "blobfile", "imageio"):
        if name in sys.modules:
            continue
        try:
            if importlib.util.find_spec(name) is not None:
                continue
        except (ImportError, ValueError):
            pass
        if stubs not in sys.path:
            sys.path.append(stubs)


def load_reference():
    """The reference package under the alias (hot-path modules: unet, rpe, nn, gaussian_diffusion, respace, script_util, ...)."""
    if ALIAS in sys.modules:
        return sys.modules[ALIAS]
    ref = require_ref()
    pkg_dir = os.path.join(ref, "improved_diffusion")
    spec = importlib.util.spec_from_file_location(ALIAS, os.path.join(pkg_dir, "__init__.py"), submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[ALIAS] = mod
    spec.loader.exec_module(mod)
    return mod


def ref_module(name):
    load_reference()
    return importlib.import_module(f"{ALIAS}.{name}")


def create_reference(over, pixel=True):
    """(model, diffusion) built by the reference's own create_model_and_diffusion with its own defaults + `over`."""
    su = ref_module("script_util")
    d = su.model_and_diffusion_defaults()
    d.update(over)
    if pixel:
        d["diffusion_space_kwargs"] = dict(diffusion_space="pixel", pre_encoded=False, pre_encoded_stats_dict=None)
    return su.create_model_and_diffusion(**d)


def enable_mixed_package():
    """Hot path from this repo, the rest of `improved_diffusion` from the reference (call BEFORE importing improved_diffusion,
    or it patches the already-imported package's search path)."""
    ref = require_ref()
    add_stubs()
    os.environ["FDM_REFERENCE_PATH"] = ref
    pkg = sys.modules.get("improved_diffusion")
    if pkg is not None:
        cand = os.path.join(ref, "improved_diffusion")
        if cand not in pkg.__path__:
            pkg.__path__.append(cand)
    return ref


def load_script(name):
    """scripts/<name>.py of the reference as a module; its `from improved_diffusion import ...` lines resolve through the mixed
    package (enable_mixed_package is called here)."""
    ref = enable_mixed_package()
    key = f"fdm_ref_script_{name}"
    if key in sys.modules:
        return sys.modules[key]
    spec = importlib.util.spec_from_file_location(key, os.path.join(ref, "scripts", f"{name}.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[key] = mod
    spec.loader.exec_module(mod)
    return mod
