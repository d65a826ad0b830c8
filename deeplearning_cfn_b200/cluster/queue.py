"""Message queues with SQS semantics - the rendezvous medium of the reference.

The reference's agents and Lambda talk through two SQS queues (deeplearning.template:743-754): the
*master queue* (Lambda -> master agent, `asg-setup` messages, deleted after processing,
dl_cfn_setup.py:122-167) and the *worker queue* (master agent -> all workers, one `worker-setup`
message read with visibility_timeout=0 and never deleted so that every worker sees it - the
broadcast-over-queue trick, dl_cfn_setup.py:169-207).  What the rendezvous logic relies on is exactly:

  * send(body)                                           at-least-once delivery (duplicates possible)
  * receive(max_messages, visibility_timeout) -> msgs    a received message is hidden for the timeout
  * delete(receipt)                                      only then is it gone

`MemoryQueue` (threads, unit tests) and `FileQueue` (a spool directory shared by the processes of one
box - the B200 stand-in for SQS) implement that contract; both can inject duplicates to exercise the
de-duplication path (dl_cfn_setup.py:141-148).
"""
from __future__ import annotations

import json
import os
import threading
import time
import uuid
from dataclasses import dataclass


@dataclass
class Message:
    body: str
    message_id: str
    receipt: str

    def json(self):
        return json.loads(self.body)


class QueueBase:
    name: str

    def send(self, body) -> str:
        raise NotImplementedError

    def receive(self, max_messages: int = 10, visibility_timeout: float = 60.0) -> list[Message]:
        raise NotImplementedError

    def delete(self, msg: Message) -> None:
        raise NotImplementedError

    def purge(self) -> None:
        raise NotImplementedError

    @staticmethod
    def _encode(body) -> str:
        return body if isinstance(body, str) else json.dumps(body)


class MemoryQueue(QueueBase):
    def __init__(self, name: str = "queue", duplicate_every: int = 0):
        self.name = name
        self._lock = threading.Lock()
        self._msgs: dict[str, dict] = {}
        self._sent = 0
        self.duplicate_every = duplicate_every

    def send(self, body) -> str:
        body = self._encode(body)
        with self._lock:
            self._sent += 1
            copies = 2 if self.duplicate_every and self._sent % self.duplicate_every == 0 else 1
            mid = ""
            for _ in range(copies):
                mid = uuid.uuid4().hex
                self._msgs[mid] = {"body": body, "visible_at": 0.0, "seq": time.monotonic()}
        return mid

    def receive(self, max_messages: int = 10, visibility_timeout: float = 60.0) -> list[Message]:
        now = time.monotonic()
        out = []
        with self._lock:
            for mid, m in sorted(self._msgs.items(), key=lambda kv: kv[1]["seq"]):
                if m["visible_at"] <= now:
                    m["visible_at"] = now + visibility_timeout
                    out.append(Message(m["body"], mid, mid))
                    if len(out) >= max_messages:
                        break
        return out

    def delete(self, msg: Message) -> None:
        with self._lock:
            self._msgs.pop(msg.receipt, None)

    def purge(self) -> None:
        with self._lock:
            self._msgs.clear()

    def __len__(self):
        with self._lock:
            return len(self._msgs)


class FileQueue(QueueBase):
    """Spool directory: one JSON file per message; visibility is a timestamp inside the file, updated
    under an O_EXCL lock file so that concurrent receivers on the same box do not double-claim."""

    def __init__(self, path: str, duplicate_every: int = 0):
        self.path = path
        self.name = os.path.basename(path.rstrip("/"))
        self.duplicate_every = duplicate_every
        os.makedirs(path, exist_ok=True)

    # -- tiny directory lock -------------------------------------------------
    def _lock(self):
        lock = os.path.join(self.path, ".lock")
        deadline = time.monotonic() + 10.0
        while True:
            try:
                fd = os.open(lock, os.O_CREAT | os.O_EXCL | os.O_WRONLY)
                os.close(fd)
                return lock
            except FileExistsError:
                if time.monotonic() > deadline:  # stale lock from a killed process
                    try:
                        os.unlink(lock)
                    except FileNotFoundError:
                        pass
                    deadline = time.monotonic() + 10.0
                time.sleep(0.002)

    def send(self, body) -> str:
        body = self._encode(body)
        lock = self._lock()
        try:
            n = len([f for f in os.listdir(self.path) if f.endswith(".msg")]) + 1
            counter = os.path.join(self.path, ".sent")
            sent = int(open(counter).read() or 0) + 1 if os.path.exists(counter) else 1
            open(counter, "w").write(str(sent))
            copies = 2 if self.duplicate_every and sent % self.duplicate_every == 0 else 1
            mid = ""
            for _ in range(copies):
                mid = f"{time.time_ns():020d}-{uuid.uuid4().hex[:8]}"
                tmp = os.path.join(self.path, mid + ".tmp")
                with open(tmp, "w") as f:
                    json.dump({"body": body, "visible_at": 0.0}, f)
                os.rename(tmp, os.path.join(self.path, mid + ".msg"))
            del n
        finally:
            os.unlink(lock)
        return mid

    def receive(self, max_messages: int = 10, visibility_timeout: float = 60.0) -> list[Message]:
        out = []
        lock = self._lock()
        try:
            now = time.time()
            for fn in sorted(f for f in os.listdir(self.path) if f.endswith(".msg")):
                full = os.path.join(self.path, fn)
                try:
                    with open(full) as f:
                        m = json.load(f)
                except (FileNotFoundError, json.JSONDecodeError):
                    continue
                if m["visible_at"] <= now:
                    if visibility_timeout > 0:
                        m["visible_at"] = now + visibility_timeout
                        tmp = full + ".tmp"
                        with open(tmp, "w") as f:
                            json.dump(m, f)
                        os.rename(tmp, full)
                    out.append(Message(m["body"], fn[:-4], fn))
                    if len(out) >= max_messages:
                        break
        finally:
            os.unlink(lock)
        return out

    def delete(self, msg: Message) -> None:
        try:
            os.unlink(os.path.join(self.path, msg.receipt))
        except FileNotFoundError:
            pass

    def purge(self) -> None:
        for fn in os.listdir(self.path):
            if fn.endswith(".msg"):
                try:
                    os.unlink(os.path.join(self.path, fn))
                except FileNotFoundError:
                    pass

    def __len__(self):
        return len([f for f in os.listdir(self.path) if f.endswith(".msg")])


def open_queue(url: str) -> QueueBase:
    """`file:///path` or a plain path -> FileQueue; `mem://name` -> process-local MemoryQueue."""
    if url.startswith("mem://"):
        return _MEM.setdefault(url, MemoryQueue(url[6:]))
    if url.startswith("file://"):
        url = url[7:]
    return FileQueue(url)


_MEM: dict[str, MemoryQueue] = {}
