"""The streaming worker: the reference's Worker.processWorker / processMessage loop
(internal/worker/worker.go:112-149,165-234) over the GPU processor.

The reference starts WORKER_CONCURRENCY goroutines (worker.go:90-96; default 3, .env.example:38); each takes one
message at a time and blocks in processor.Process: image.Decode -> operations -> encode -> SaveProcessed.  The GPU
drop-in keeps exactly that shape -- N host threads, each `decode -> Process -> (encode, save inside Process)` -- and
gets its batching from the library: Process submits the image's raster work (ipg_submit) and blocks in ipg_wait, so
while one thread waits its GPU ticket, others decode or encode, and the engine's batcher coalesces the tickets of
all threads into batched launches.  Nothing in the loop is a barrier: decode, H2D, kernels, D2H and encode of
different messages overlap.

Kafka, MinIO and Postgres are bypassed (no network here): messages come from an iterable, bytes from a callable,
objects go to the processor's file repository.  Per-stage times are recorded per message so a harness can report
decode / raster (submit -> wait: H2D + kernels + D2H) / encode / save separately, as north_star asks.
"""
from __future__ import annotations

import queue
import threading
import time
from dataclasses import dataclass, field
from typing import Callable, Iterable, List, Optional, Tuple

from . import codecs
from .processor import ImageProcessor


@dataclass
class MessageResult:
    task_id: str
    result: Optional[dict]
    error: Optional[str]
    decode_s: float = 0.0
    process_s: float = 0.0    # Process() wall time: parameter handling + submit + wait + encode + save
    encode_s: float = 0.0     # of which inside the encode callback(s)
    save_s: float = 0.0       # ... and inside SaveProcessed
    pixels: int = 0
    src_bytes: int = 0

    @property
    def raster_s(self) -> float:
        """submit -> wait of the raster work (H2D + kernels + D2H incl. queueing behind other messages)."""
        return max(self.process_s - self.encode_s - self.save_s, 0.0)


@dataclass
class WorkerStats:
    wall_s: float = 0.0
    messages: int = 0
    failed: int = 0
    decode_s: float = 0.0
    raster_s: float = 0.0
    encode_s: float = 0.0
    save_s: float = 0.0
    pixels: int = 0
    results: List[MessageResult] = field(default_factory=list)

    def summary(self) -> dict:
        n = max(self.messages, 1)
        return {"messages": self.messages, "failed": self.failed, "wall_s": self.wall_s,
                "images_per_s": self.messages / self.wall_s if self.wall_s > 0 else None,
                "megapixels_per_s": self.pixels / 1e6 / self.wall_s if self.wall_s > 0 else None,
                "thread_seconds": {"decode": self.decode_s, "raster_submit_to_wait": self.raster_s,
                                   "encode": self.encode_s, "save": self.save_s},
                "mean_ms_per_image": {"decode": 1e3 * self.decode_s / n, "raster_submit_to_wait": 1e3 * self.raster_s / n,
                                      "encode": 1e3 * self.encode_s / n, "save": 1e3 * self.save_s / n}}


class StreamingWorker:
    """`concurrency` threads over one ImageProcessor (itself over one Engine that may drive several GPUs)."""

    def __init__(self, processor: ImageProcessor, concurrency: int = 3,
                 decode: Callable[[bytes], Tuple[object, str]] = codecs.decode):
        self.processor = processor
        self.concurrency = max(1, int(concurrency))
        self.decode = decode
        self._tls = threading.local()
        # per-thread accounting of the time spent inside the processor's encode / save callbacks
        proc = processor
        inner_encode, inner_save = proc.encode, proc.file_repo.save_processed

        def timed_encode(rgba, fmt, quality):
            t0 = time.perf_counter()
            try:
                return inner_encode(rgba, fmt, quality)
            finally:
                self._tls.encode_s = getattr(self._tls, "encode_s", 0.0) + time.perf_counter() - t0

        def timed_save(path, data, ctype):
            t0 = time.perf_counter()
            try:
                return inner_save(path, data, ctype)
            finally:
                self._tls.save_s = getattr(self._tls, "save_s", 0.0) + time.perf_counter() - t0

        proc.encode = timed_encode
        proc.file_repo.save_processed = timed_save

    # worker.go:165-234 for one message
    def process_message(self, task_json: str | dict, image_bytes: bytes) -> MessageResult:
        tid = task_json.get("ID", "") if isinstance(task_json, dict) else ""
        self._tls.encode_s = self._tls.save_s = 0.0
        t0 = time.perf_counter()
        img, fmt, derr = None, "", None
        try:
            img, fmt = self.decode(image_bytes)
        except Exception as e:  # noqa: BLE001 -- image.Decode error text goes into the result (image_processor.go:47-53)
            derr = f"{e}"
        t1 = time.perf_counter()
        res, err = self.processor.process(task_json, img, fmt or "jpeg", decode_error=derr)
        t2 = time.perf_counter()
        return MessageResult(tid or (res or {}).get("ID", ""), res, err, decode_s=t1 - t0, process_s=t2 - t1,
                             encode_s=self._tls.encode_s, save_s=self._tls.save_s,
                             pixels=(img.width * img.height) if img is not None else 0, src_bytes=len(image_bytes))

    # worker.go:112-149: N goroutines draining the message source
    def run(self, messages: Iterable[Tuple[str | dict, bytes]]) -> WorkerStats:
        q: "queue.Queue" = queue.Queue(maxsize=4 * self.concurrency)
        out: List[Optional[MessageResult]] = []
        lock = threading.Lock()

        def loop():
            while True:
                item = q.get()
                if item is None:
                    return
                idx, task, data = item
                r = self.process_message(task, data)
                with lock:
                    out[idx] = r

        threads = [threading.Thread(target=loop, daemon=True) for _ in range(self.concurrency)]
        t0 = time.perf_counter()
        for t in threads:
            t.start()
        n = 0
        for task, data in messages:
            with lock:
                out.append(None)
            q.put((n, task, data))
            n += 1
        for _ in threads:
            q.put(None)
        for t in threads:
            t.join()
        st = WorkerStats(wall_s=time.perf_counter() - t0, messages=n)
        for r in out:
            st.results.append(r)
            st.failed += 1 if (r is None or r.error) else 0
            if r is not None:
                st.decode_s += r.decode_s
                st.raster_s += r.raster_s
                st.encode_s += r.encode_s
                st.save_s += r.save_s
                st.pixels += r.pixels
        return st
