from .context import get_context_scheduler  # noqa: F401
