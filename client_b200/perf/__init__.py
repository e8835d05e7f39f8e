"""client_b200.perf -- device-side load generation (perf_analyzer-style CLI in
``client_b200.perf.cli``; ``python -m client_b200.perf``)."""

from .loadgen import ConcurrencyManager, InferStat, RequestRecord, SlotSet, TensorSpec, measure, summarize

__all__ = ["ConcurrencyManager", "InferStat", "RequestRecord", "SlotSet", "TensorSpec", "measure", "summarize"]
