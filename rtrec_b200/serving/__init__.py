"""Serving callers of the SLIM hot path: the FastAPI app and a micro-batching stream consumer."""
