"""FastAPI front of the device-resident SLIM model.

Same routes, request/response models, token check and status codes as the reference's app
(/root/reference/rtrec/serving/app.py:35-93; behaviour pinned by /root/reference/tests/serving/test_app.py:48-109):

    GET  /           -> {"message": "Recommender System API is running"}
    POST /fit        -> List[Interaction]  (X-Token header)  -> {"message": "Training successful"}
    POST /recommend  -> RecommendationRequest (X-Token)      -> {"user": ..., "recommendations": [...]}

What differs is below the routes.  The reference calls ``SLIM.fit`` / ``SLIM.recommend`` straight from the event loop;
here the model owns one CUDA stream and grow-only device scratch (include/rtrec_b200.h "Threading": one caller at a
time), so every model call goes through one lock and runs on a worker thread (``run_in_threadpool``): a fit in progress
never blocks the event loop, and requests are applied in arrival order.  ``StreamIngestor`` is the same discipline for
record streams (the reference's Kinesis consumer, /root/reference/examples/kinesis/kinesis_consumer.py:87-99: a training
lock around ``recommender.fit`` per record): records are parsed as they arrive and folded + re-solved once per
micro-batch, because one partial fit of 1,000 events costs the device about as much as one of a single event.
"""
from __future__ import annotations

import json
import logging
import os
import threading
import time
from typing import Any, Callable, Iterable, List, Optional, Tuple

from fastapi import FastAPI, Header, HTTPException
from fastapi.concurrency import run_in_threadpool
from fastapi.middleware.cors import CORSMiddleware
from pydantic import BaseModel

from ..models import SLIM

DEFAULT_SECRET_TOKEN = "fake_secret_token"
SECRET_TOKEN = os.getenv("X_TOKEN", DEFAULT_SECRET_TOKEN)

logging.basicConfig(level=logging.INFO)


class Interaction(BaseModel):
    user: Any
    item: Any
    timestamp: float
    rating: float


class RecommendationRequest(BaseModel):
    user: Any
    top_k: int = 10
    filter_interacted: bool = True


class RecommendationResponse(BaseModel):
    user: Any
    recommendations: List[Any]


class ModelGate:
    """One model, one caller at a time (the device library is single-caller and stream-ordered)."""

    def __init__(self, model: SLIM):
        self.model = model
        self.lock = threading.Lock()

    def fit(self, events: List[Tuple[Any, Any, float, float]], update_interaction: bool = False) -> None:
        with self.lock:
            self.model.fit(events, update_interaction=update_interaction, progress_bar=False)

    def recommend(self, user: Any, top_k: int, filter_interacted: bool) -> List[Any]:
        with self.lock:
            return self.model.recommend(user=user, top_k=top_k, filter_interacted=filter_interacted)


class StreamIngestor:
    """Micro-batching consumer for record streams (the role of kinesis_consumer.py:87-99 ``process_records`` /
    ``run_task``).  ``put(records)`` may be called from any thread or task; records are JSON objects (or bytes / str
    holding one) with ``user, item, timestamp, rating``.  A background thread folds whatever has accumulated -- at most
    ``max_batch`` events, at least every ``max_wait_s`` seconds -- with ONE partial fit under the model lock.
    Malformed records are logged and dropped, like base.py:85-94 does per event."""

    def __init__(self, gate: ModelGate, max_batch: int = 65536, max_wait_s: float = 0.25, update_interaction: bool = False,
                 on_batch: Optional[Callable[[int, float], None]] = None):
        self.gate, self.max_batch, self.max_wait_s = gate, int(max_batch), float(max_wait_s)
        self.update_interaction = update_interaction
        self.on_batch = on_batch
        self._buf: List[Tuple[Any, Any, float, float]] = []
        self._cv = threading.Condition()
        self._stop = False
        self.n_events = 0
        self.n_batches = 0
        self._thread = threading.Thread(target=self._run, name="rtrec-ingest", daemon=True)
        self._thread.start()

    @staticmethod
    def parse(record: Any) -> Tuple[Any, Any, float, float]:
        if isinstance(record, (bytes, bytearray)):
            record = record.decode()
        if isinstance(record, str):
            record = json.loads(record)
        if isinstance(record, dict) and "Data" in record:        # a raw Kinesis record
            return StreamIngestor.parse(record["Data"])
        return record["user"], record["item"], float(record["timestamp"]), float(record["rating"])

    def put(self, records: Iterable[Any]) -> int:
        events = []
        for rec in records:
            try:
                events.append(self.parse(rec))
            except Exception as e:  # noqa: BLE001
                logging.warning(f"Error processing record: {e}")
        with self._cv:
            self._buf.extend(events)
            self._cv.notify()
        return len(events)

    def _run(self) -> None:
        while True:
            with self._cv:
                if not self._buf and not self._stop:
                    self._cv.wait(self.max_wait_s)
                if self._stop and not self._buf:
                    return
                batch, self._buf = self._buf[:self.max_batch], self._buf[self.max_batch:]
            if not batch:
                continue
            t0 = time.perf_counter()
            try:
                self.gate.fit(batch, update_interaction=self.update_interaction)
            except Exception as e:  # noqa: BLE001
                logging.error(f"Training failed: {e}")
            self.n_events += len(batch)
            self.n_batches += 1
            if self.on_batch:
                self.on_batch(len(batch), time.perf_counter() - t0)

    def flush(self, timeout: float = 30.0) -> None:
        """Block until everything put so far has been trained on."""
        end = time.time() + timeout
        while time.time() < end:
            with self._cv:
                empty = not self._buf
            if empty and not self.gate.lock.locked():
                with self.gate.lock:      # a fit that took the batch before we looked has finished
                    pass
                with self._cv:
                    if not self._buf:
                        return
            time.sleep(0.005)
        raise TimeoutError("StreamIngestor.flush timed out")

    def close(self) -> None:
        with self._cv:
            self._stop = True
            self._cv.notify()
        self._thread.join(timeout=30)


def create_app(model: Optional[SLIM] = None) -> FastAPI:
    """Factory (app.py:35-93).  ``model``: serve an existing (e.g. loaded) model instead of a fresh one."""
    app = FastAPI()
    app.add_middleware(CORSMiddleware, allow_origins=["*"], allow_credentials=True, allow_methods=["*"], allow_headers=["*"])
    gate = ModelGate(model if model is not None else SLIM(min_value=-5, max_value=10, decay_in_days=365))
    app.state.gate = gate

    @app.get("/")
    def read_root():
        return {"message": "Recommender System API is running"}

    @app.post("/fit")
    async def fit(interactions: List[Interaction], x_token: str = Header()):
        if x_token != SECRET_TOKEN:
            raise HTTPException(status_code=400, detail="Invalid X-Token header")
        try:
            events = [(it.user, it.item, it.timestamp, it.rating) for it in interactions]
            await run_in_threadpool(gate.fit, events)
            return {"message": "Training successful"}
        except Exception as e:  # noqa: BLE001
            logging.error(f"Training failed: {e}")
            raise HTTPException(status_code=500, detail="Training failed")

    @app.post("/recommend", response_model=RecommendationResponse)
    async def recommend(request: RecommendationRequest, x_token: str = Header()):
        if x_token != SECRET_TOKEN:
            raise HTTPException(status_code=400, detail="Invalid X-Token header")
        try:
            recs = await run_in_threadpool(gate.recommend, request.user, request.top_k, request.filter_interacted)
            return {"user": request.user, "recommendations": recs}
        except Exception as e:  # noqa: BLE001
            logging.error(f"Recommendation failed: {e}")
            raise HTTPException(status_code=500, detail="Recommendation failed")

    return app


if __name__ == "__main__":
    import uvicorn
    uvicorn.run(create_app(), host="0.0.0.0", port=8000)
