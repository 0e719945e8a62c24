"""Result containers of the tool boundary (reference: fenics_mcp_server.py:168-197).

Same attribute set as the reference dataclasses; defined in an importable module so pickles
written by the server load anywhere this package is importable."""
from dataclasses import dataclass
from typing import Any, Dict, List


@dataclass
class TimeSeriesField:
    """Scalar field time series: coords [N][3], values [Nt][N], times [Nt], dim, meta."""
    coords: List[List[float]]
    values: List[List[float]]
    times: List[float]
    dim: int
    meta: Dict[str, Any]


@dataclass
class SolveResult:
    """Path of the pickled TimeSeriesField plus metadata."""
    data_file: str
    dim: int
    meta: Dict[str, Any]


@dataclass
class PlotResult:
    html_path: str
