from dataclasses import dataclass
from typing import Any


@dataclass
class WanPipelineOutput:
    frames: Any
