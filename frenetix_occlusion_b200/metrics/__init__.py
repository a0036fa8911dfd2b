"""Pluggable metrics of the assessment path (mirrors the reference's ``frenetix_occlusion/metrics``
package: same class names, constructor arguments and ``evaluate(trajectory, results)`` protocol).
All numbers come from the CUDA dense core (``engine.MetricEngine``)."""
from .metric import Metric  # noqa: F401
from .cp import CP  # noqa: F401
from .dce import DCE  # noqa: F401
from .ttc import TTC  # noqa: F401
from .ttce import TTCE  # noqa: F401
from .wttc import WTTC  # noqa: F401
from .be import BE  # noqa: F401
from .hr import HR  # noqa: F401
