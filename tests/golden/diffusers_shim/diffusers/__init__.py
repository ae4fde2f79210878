"""Minimal stand-in for the `diffusers` package, used ONLY by tests/golden/make_golden.py to import and execute the
reference's own source files (/root/reference/architecture/*.py) in a container where diffusers is not installed.

It supplies just the upstream symbols those files — and, for the pipeline golden, the reference's
pipelines/pipeline_wan_i2v_motion_FrameINO.py — import. Anything with arithmetic (FeedForward, FP32LayerNorm,
RMSNorm, AdaLayerNorm, CogVideoXLayerNormZero, activations) is restated from recalled upstream diffusers (~0.33-0.35)
semantics; everything else is inert plumbing. Not imported by the product or by the tests at run time.
"""
__version__ = "0.35.0.shim"
