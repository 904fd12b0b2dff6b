"""ComfyUI entry point: dropping this repository into ComfyUI/custom_nodes/ registers the MixDQ
nodes, exactly like the reference plugin's repo-root __init__.py (/root/reference/__init__.py:1-3,
which re-exports kernels/mixdq.py:779-791). Outside ComfyUI (tests, bench) this file is not
imported: `mixdq_b200` and `mixdq_extension` are used as top-level packages."""
try:
    from .mixdq_b200.mixdq import NODE_CLASS_MAPPINGS, NODE_DISPLAY_NAME_MAPPINGS
except ImportError:      # imported as a plain module (no parent package), e.g. by a test runner
    from mixdq_b200.mixdq import NODE_CLASS_MAPPINGS, NODE_DISPLAY_NAME_MAPPINGS

__all__ = ["NODE_CLASS_MAPPINGS", "NODE_DISPLAY_NAME_MAPPINGS"]
