"""Minimal stand-in for ``diffusers==0.29.2`` (TEST INFRASTRUCTURE ONLY, never shipped).

``diffusers`` is not installable offline, and every hot-path file of the reference imports it.
This shim restates just the leaves the reference touches so that
``/root/reference/modules/*.py`` and ``pipelines/v_express_pipeline.py`` can be imported and run
*verbatim* by ``oracle/gen_golden.py`` to produce the committed golden vectors.
"""
from .pipeline_utils import DiffusionPipeline  # noqa: F401
from .schedulers import DDIMScheduler  # noqa: F401
from .autoencoder import AutoencoderKL  # noqa: F401
