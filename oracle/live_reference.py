"""Import the UNMODIFIED reference live: from /root/reference in the build container, from the byte-for-byte staged copy
under baseline/_ref/ (oracle/stage_ref.py; git-ignored) on the GPU box.

Test infrastructure.  Nothing is copied into the repo: ``modules.py`` / ``networks.py`` are
imported by path; ``model.py`` is read as text and exec'd with the single reporting shim
``.data[0] -> .item()`` because 0-dim indexing raises on torch>=0.5.
``available()`` is False when neither exists and callers must skip.
"""
import os
import sys
import types
import warnings

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _resolve():
    for d in (os.environ.get("DTG_REFERENCE_DIR"), "/root/reference/augmented_cyclegan",
              os.path.join(_ROOT, "baseline", "_ref", "augmented_cyclegan")):
        if d and os.path.isfile(os.path.join(d, "model.py")):
            return d
    return "/root/reference/augmented_cyclegan"


REF_DIR = _resolve()


def available():
    return os.path.isfile(os.path.join(REF_DIR, "model.py"))


_cache = {}


def load():
    """Returns (modules, networks, model) reference python modules."""
    if _cache:
        return _cache["m"], _cache["n"], _cache["M"]
    if not available():
        raise RuntimeError("reference not present at %s" % REF_DIR)
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF_DIR)
    try:
        for name in ("modules", "networks", "model"):
            sys.modules.pop(name, None)
        import modules as ref_modules      # noqa
        import networks as ref_networks    # noqa
        src = open(os.path.join(REF_DIR, "model.py")).read().replace(".data[0]", ".item()")
        ref_model = types.ModuleType("ref_model")
        ref_model.__file__ = os.path.join(REF_DIR, "model.py")
        exec(compile(src, ref_model.__file__, "exec"), ref_model.__dict__)
    finally:
        sys.path.remove(REF_DIR)
        for name in ("modules", "networks"):
            sys.modules.pop(name, None)
    _cache.update(m=ref_modules, n=ref_networks, M=ref_model)
    return ref_modules, ref_networks, ref_model


def build_reference_model(opt, state, stoch=False, ignore_noise=False):
    """AugmentedCycleGAN(opt, testing=True) -- or StochCycleGAN(opt, ignore_noise, testing=True) when
    `stoch` -- with `state` (oracle.nets.init_model_state layout) loaded."""
    _, _, M = load()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if stoch:
            model = M.StochCycleGAN(opt, ignore_noise=ignore_noise, testing=True)
        else:
            model = M.AugmentedCycleGAN(opt, testing=True)
    for name, sd in state.items():
        if not hasattr(model, name):        # StochCycleGAN has no netE_B / netD_z_B
            continue
        net = getattr(model, name)
        full = net.state_dict()
        new = {}
        for k in full:
            src = k
            if k not in sd:
                # alias keys model.1x.{1,4,5}.* -> model.1x.conv_block.{1,4,5}.* (modules.py:145-146)
                parts = k.split(".")
                src = ".".join(parts[:2] + ["conv_block"] + parts[2:])
            new[k] = sd[src].detach().clone()
        net.load_state_dict(new)
    return model
