"""Predict-side input pipeline and trainer shell behind the reference's `pdp.factorgraph` names."""
from .dataset import DynamicBatchDivider, FactorGraphDataset   # noqa: F401
from .base import FactorGraphTrainerBase                       # noqa: F401
