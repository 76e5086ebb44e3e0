from .audio import AudioPreprocessor  # noqa: F401
