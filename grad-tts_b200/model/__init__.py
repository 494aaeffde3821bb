# mirrors /root/reference/model/__init__.py:9
from .tts import GradTTS  # noqa: F401
