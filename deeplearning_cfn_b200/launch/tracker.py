"""Job trackers: the rendezvous services a launcher starts before fanning processes out.

Reference: `DT/tracker.py` - `PSTracker` (:336-386: pick a free port in 9091..9999, start the ps-lite
*scheduler* locally with DMLC_ROLE=scheduler / DMLC_PS_ROOT_URI / DMLC_PS_ROOT_PORT) and `RabitTracker`
(:137-334: socket rendezvous for all-reduce jobs - magic 0xff99 handshake :58-71, ranks assigned in
host-sorted order :295-308, binary-tree + ring link maps :185-252, commands start / recover / shutdown /
print :266-320); `submit()` (:410-433) picks Rabit when there are no servers, PS otherwise.

This is an independent implementation of the same protocol idea.  What an all-reduce worker gets back
from the tracker here is what a B200 rank needs: rank, world size, its tree parent/children and ring
neighbours (kept for API parity; on NVSwitch every peer is one hop), and the address of rank 0's
torch.distributed store so that `hvd.init()` can call `init_process_group` without any other env.
"""
from __future__ import annotations

import logging
import os
import socket
import struct
import subprocess
import threading
import time

MAGIC = 0xFF99
LOG = logging.getLogger("dlcfn-tracker")


# --------------------------------------------------------------------------- wire helpers
class Wire:
    def __init__(self, sock: socket.socket):
        self.sock = sock

    def recv_all(self, n: int) -> bytes:
        buf = b""
        while len(buf) < n:
            chunk = self.sock.recv(n - len(buf))
            if not chunk:
                raise ConnectionError("peer closed")
            buf += chunk
        return buf

    def recv_int(self) -> int:
        return struct.unpack("@i", self.recv_all(4))[0]

    def send_int(self, v: int) -> None:
        self.sock.sendall(struct.pack("@i", v))

    def recv_str(self) -> str:
        return self.recv_all(self.recv_int()).decode()

    def send_str(self, s: str) -> None:
        b = s.encode()
        self.send_int(len(b))
        self.sock.sendall(b)


# --------------------------------------------------------------------------- topology maps
def tree_neighbors(r: int, n: int) -> list[int]:
    """Binary-heap tree: parent (r+1)//2-1, children 2r+1, 2r+2."""
    out = []
    if r > 0:
        out.append((r + 1) // 2 - 1)
    for c in (2 * r + 1, 2 * r + 2):
        if c < n:
            out.append(c)
    return out


def share_ring(n: int) -> list[int]:
    """A Hamiltonian order over the heap tree that reuses tree edges where it can: depth-first walk, the
    last subtree of every node reversed so the walk comes back next to its parent."""
    def walk(r: int) -> list[int]:
        kids = [c for c in (2 * r + 1, 2 * r + 2) if c < n]
        order = [r]
        for i, c in enumerate(kids):
            sub = walk(c)
            if i == len(kids) - 1:
                sub.reverse()
            order += sub
        return order
    return walk(0) if n > 0 else []


def link_maps(n: int):
    """(tree_map, parent_map, ring_map) with ranks relabelled so that the ring is 0 -> 1 -> ... -> n-1."""
    order = share_ring(n)
    relabel = {old: new for new, old in enumerate(order)}
    tree_map, parent_map, ring_map = {}, {}, {}
    for old in range(n):
        new = relabel[old]
        tree_map[new] = [relabel[x] for x in tree_neighbors(old, n)]
        parent_map[new] = relabel[(old + 1) // 2 - 1] if old > 0 else -1
        ring_map[new] = ((new - 1) % n, (new + 1) % n)
    return tree_map, parent_map, ring_map


# --------------------------------------------------------------------------- all-reduce tracker
def _free_port(host: str, lo: int, hi: int) -> tuple[socket.socket, int]:
    for port in range(lo, hi):
        s = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        s.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        try:
            s.bind((host, port))
            return s, port
        except OSError:
            s.close()
    raise OSError(f"no free port in [{lo}, {hi})")


class AllreduceTracker:
    """Rendezvous for `nworker` all-reduce ranks (the RabitTracker role)."""

    def __init__(self, host_ip: str, nworker: int, port: int = 9091, port_end: int = 9999):
        self.host_ip, self.nworker = host_ip, nworker
        self.sock, self.port = _free_port(host_ip, port, port_end)
        self.sock.listen(256)
        self.thread: threading.Thread | None = None
        self.messages: list[str] = []
        self.start_time = self.end_time = None
        self.error: Exception | None = None

    def slave_envs(self) -> dict:
        return {"DMLC_TRACKER_URI": self.host_ip, "DMLC_TRACKER_PORT": self.port}

    def _serve(self) -> None:
        n = self.nworker
        tree_map, parent_map, ring_map = link_maps(n)
        job_rank: dict[str, int] = {}
        pending: list = []          # (host, wire, jobid, store_port) waiting for a fresh rank
        shutdown: set[int] = set()
        master = None               # (host, port) of rank 0's store
        assigned: dict[int, tuple] = {}
        todo = list(range(n))

        def send_assignment(w: Wire, rank: int):
            w.send_int(rank)
            w.send_int(parent_map[rank])
            w.send_int(n)
            nb = tree_map[rank]
            w.send_int(len(nb))
            for x in nb:
                w.send_int(x)
            prev, nxt = ring_map[rank]
            w.send_int(prev)
            w.send_int(nxt)
            w.send_str(master[0])
            w.send_int(master[1])

        while len(shutdown) != n:
            conn, addr = self.sock.accept()
            w = Wire(conn)
            try:
                if w.recv_int() != MAGIC:
                    conn.close()
                    continue
                w.send_int(MAGIC)
                rank, world, jobid, cmd = w.recv_int(), w.recv_int(), w.recv_str(), w.recv_str()
            except (ConnectionError, struct.error):
                conn.close()
                continue
            if cmd == "print":
                self.messages.append(w.recv_str())
                LOG.info(self.messages[-1].strip())
                conn.close()
                continue
            if cmd == "shutdown":
                shutdown.add(rank)
                conn.close()
                continue
            if cmd not in ("start", "recover"):
                conn.close()
                continue
            if world not in (-1, n):
                conn.close()
                continue
            store_port = w.recv_int()   # a free port on the worker's host, used if it becomes rank 0
            host = addr[0]
            if cmd == "recover" or (rank == -1 and jobid != "NULL" and jobid in job_rank):
                r = rank if rank >= 0 else job_rank[jobid]
                if master is not None and r in assigned:
                    send_assignment(w, r)
                    conn.close()
                    continue
            pending.append((host, w, jobid, store_port))
            if len(pending) == len(todo):
                # ranks in host-sorted order, like the reference (:295-308)
                pending.sort(key=lambda p: p[0])
                if self.start_time is None:
                    self.start_time = time.time()
                for (h, pw, jid, sp), r in zip(pending, todo):
                    assigned[r] = (h, sp)
                    if jid != "NULL":
                        job_rank[jid] = r
                master = assigned[0]
                for (h, pw, jid, sp), r in zip(pending, todo):
                    try:
                        send_assignment(pw, r)
                    finally:
                        pw.sock.close()
                pending, todo = [], []
        self.end_time = time.time()

    def start(self) -> None:
        def run():
            try:
                self._serve()
            except Exception as e:   # surfaced by join()
                self.error = e
        self.thread = threading.Thread(target=run, daemon=True, name="allreduce-tracker")
        self.thread.start()

    def join(self, timeout: float | None = None) -> None:
        t_end = None if timeout is None else time.time() + timeout
        while self.thread is not None and self.thread.is_alive():
            self.thread.join(0.1)
            if t_end is not None and time.time() > t_end:
                break
        if self.error:
            raise self.error

    def alive(self) -> bool:
        return self.thread is not None and self.thread.is_alive()

    def stop(self) -> None:
        """End the rendezvous service (the launcher calls this when a rank died: nobody will send its `shutdown`)."""
        try:
            self.sock.close()          # accept() in the service thread fails, the thread ends (daemon)
        except OSError:
            pass


class TrackerClient:
    """Worker side of the AllreduceTracker protocol."""

    def __init__(self, uri: str | None = None, port: int | None = None):
        self.uri = uri or os.environ["DMLC_TRACKER_URI"]
        self.port = int(port or os.environ["DMLC_TRACKER_PORT"])

    def _connect(self, rank: int, world: int, jobid: str, cmd: str) -> Wire:
        s = socket.create_connection((self.uri, self.port), timeout=60)
        w = Wire(s)
        w.send_int(MAGIC)
        if w.recv_int() != MAGIC:
            raise ConnectionError("tracker handshake failed")
        w.send_int(rank)
        w.send_int(world)
        w.send_str(jobid)
        w.send_str(cmd)
        return w

    def start(self, world: int = -1, jobid: str = "NULL", rank: int = -1, cmd: str = "start") -> dict:
        probe = socket.socket()
        probe.bind(("", 0))
        store_port = probe.getsockname()[1]
        probe.close()
        w = self._connect(rank, world, jobid, cmd)
        w.send_int(store_port)
        w.sock.settimeout(600)
        out = {"rank": w.recv_int(), "parent": w.recv_int(), "world": w.recv_int()}
        out["neighbors"] = [w.recv_int() for _ in range(w.recv_int())]
        out["ring_prev"], out["ring_next"] = w.recv_int(), w.recv_int()
        out["master_addr"], out["master_port"] = w.recv_str(), w.recv_int()
        w.sock.close()
        return out

    def print(self, msg: str) -> None:
        w = self._connect(-1, -1, "NULL", "print")
        w.send_str(msg)
        w.sock.close()

    def shutdown(self, rank: int) -> None:
        self._connect(rank, -1, "NULL", "shutdown").sock.close()


# --------------------------------------------------------------------------- parameter-server tracker
class PSTracker:
    """Starts the scheduler process of a parameter-server job and owns its address."""

    def __init__(self, host_ip: str, cmd: str | None, port: int = 9091, port_end: int = 9999, envs: dict | None = None):
        self.cmd = cmd
        self.host_ip = host_ip
        self.proc: subprocess.Popen | None = None
        if cmd is None:
            self.port = None
            return
        # the reference scans 9091..9999 from the bottom; concurrent jobs on one box would all probe the same port first
        # and race between this probe and the scheduler's own bind, so start at a per-process offset inside the range
        start = port + (os.getpid() * 7919) % max(1, (port_end - port) // 2) if port_end - port > 64 else port
        probe, self.port = _free_port("", start, port_end)
        probe.close()
        env = os.environ.copy()
        env["DMLC_ROLE"] = "scheduler"
        env["DMLC_PS_ROOT_URI"] = str(host_ip)
        env["DMLC_PS_ROOT_PORT"] = str(self.port)
        for k, v in (envs or {}).items():
            env[k] = str(v)
        self.proc = subprocess.Popen(cmd, shell=True, env=env)

    def slave_envs(self) -> dict:
        if self.cmd is None:
            return {}
        return {"DMLC_PS_ROOT_URI": self.host_ip, "DMLC_PS_ROOT_PORT": self.port}

    def join(self, timeout: float | None = None) -> int | None:
        if self.proc is None:
            return None
        try:
            return self.proc.wait(timeout)
        except subprocess.TimeoutExpired:
            return None

    def alive(self) -> bool:
        return self.proc is not None and self.proc.poll() is None

    def stop(self) -> None:
        if self.alive():
            kill_tree(self.proc)


def kill_tree(proc) -> None:
    """Kill a launched process AND its descendants: the launchers start commands through a shell, so the Popen handle is
    the shell and the node itself is its child - killing only the handle would orphan a node that is blocked in a barrier."""
    if proc is None or proc.poll() is not None:
        return
    victims = []
    try:
        import psutil
        victims = psutil.Process(proc.pid).children(recursive=True)
    except Exception:       # noqa: BLE001 - psutil missing or the process just ended
        victims = []
    for v in victims:
        try:
            v.kill()
        except Exception:   # noqa: BLE001
            pass
    try:
        proc.kill()
    except OSError:
        pass


def get_host_ip(host_ip: str | None = None) -> str:
    """'auto'/'ip' -> this host's address; falls back to loopback (the container hostname may not resolve)."""
    if host_ip in (None, "auto", "ip"):
        try:
            ip = socket.gethostbyname(socket.getfqdn())
        except OSError:
            ip = "127.0.0.1"
        if ip.startswith("127."):
            try:
                s = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
                s.connect(("10.255.255.255", 1))
                ip = s.getsockname()[0]
                s.close()
            except OSError:
                ip = "127.0.0.1"
        return ip
    if host_ip == "dns":
        return socket.getfqdn()
    return host_ip


def submit(nworker: int, nserver: int, fun_submit, host_ip: str = "auto", pscmd: str | None = None,
           wait: bool = True):
    """Start the right tracker, then hand {DMLC_NUM_WORKER, DMLC_NUM_SERVER, tracker envs} to the
    launcher-specific `fun_submit(nworker, nserver, envs)` (reference :410-433)."""
    if nserver == 0:
        pscmd = None
    envs = {"DMLC_NUM_WORKER": nworker, "DMLC_NUM_SERVER": nserver}
    ip = get_host_ip(host_ip)
    if nserver == 0:
        tracker = AllreduceTracker(ip, nworker)
        envs.update(tracker.slave_envs())
        tracker.start()
    else:
        tracker = PSTracker(ip, pscmd, envs=envs)
        envs.update(tracker.slave_envs())
    handle = fun_submit(nworker, nserver, envs)

    def first_failure():
        """(name, code) of a launched process that ended with a non-zero code after its retries, else None."""
        for k, rc in list(getattr(handle, "returncodes", {}).items()):
            if rc not in (0, None):
                return k, rc
        return None

    def stop_job(why: str) -> None:
        # A synchronous job cannot make progress without one of its members (the reference documents that a dead worker
        # stalls everyone, distributed_training.md:101, and its local launcher then waits forever): end it instead.
        logging.error("%s - stopping the job", why)
        for p in list(getattr(handle, "procs", [])):
            kill_tree(p)
        tracker.stop()

    if wait:
        if isinstance(tracker, AllreduceTracker):
            # the job is over when every launched process has exited or every rank sent `shutdown`
            while tracker.alive():
                if handle is not None and hasattr(handle, "all_done") and handle.all_done():
                    break
                bad = first_failure()
                if bad:
                    stop_job(f"{bad[0]} exited with code {bad[1]}")
                    break
                time.sleep(0.1)
        else:
            rc = None
            while True:
                rc = tracker.join(timeout=0.2)
                if rc is not None or not tracker.alive():
                    break
                bad = first_failure()
                if bad:
                    stop_job(f"{bad[0]} exited with code {bad[1]}")
                    rc = tracker.join(timeout=10)
                    rc = None          # reported through the handle's return codes
                    break
            if rc not in (0, None) and handle is not None:
                # the scheduler died (could not bind, crashed): the other nodes can only spin on a dead address
                logging.error("scheduler exited with code %s - stopping the job", rc)
                for p in list(getattr(handle, "procs", [])):
                    kill_tree(p)
        if handle is not None and hasattr(handle, "wait"):
            handle.wait()
    return tracker, handle
