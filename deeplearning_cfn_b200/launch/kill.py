"""Stop a distributed job on every host of a hostfile.

Reference: `MX/tools/kill-mxnet.py <hostfile> <user> <prog>` ssh-es to each host and `kill -9`s the
processes of <user> whose command line matches <prog> (:33-55).  Same CLI; matching is done on this box
by scanning /proc (exact uid + substring of the command line), never through pkill patterns, and the
caller's own process tree is excluded.
"""
from __future__ import annotations

import os
import pwd
import signal
import sys

from .hostfile import parse_hostfile
from .transport import ssh_argv, use_local


def local_matches(user: str, prog: str) -> list[int]:
    try:
        uid = pwd.getpwnam(user).pw_uid
    except KeyError:
        return []
    me = {os.getpid(), os.getppid()}
    out = []
    for d in os.listdir("/proc"):
        if not d.isdigit() or int(d) in me:
            continue
        try:
            if os.stat(f"/proc/{d}").st_uid != uid:
                continue
            cmd = open(f"/proc/{d}/cmdline", "rb").read().replace(b"\0", b" ").decode(errors="replace")
        except OSError:
            continue
        if prog in cmd and "launch.kill" not in cmd:
            out.append(int(d))
    return out


def remote_command(user: str, prog: str) -> str:
    # list candidate pids for the user, filter on the command line, kill exactly those pids
    return (f"for p in $(ps -u {user} -o pid=); do "
            f"if tr '\\0' ' ' < /proc/$p/cmdline 2>/dev/null | grep -qF -- '{prog}'; then kill -9 $p; fi; done")


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    if len(argv) != 3:
        print("usage: kill <hostfile> <user> <prog>      (ends the processes of <user> whose command line contains <prog> on every host)")
        return 0 if argv and argv[0] in ("-h", "--help") else 1
    hostfile, user, prog = argv
    import subprocess
    done_local = False
    for h in parse_hostfile(hostfile):
        if use_local(h.name):
            if not done_local:
                for pid in local_matches(user, prog):
                    try:
                        os.kill(pid, signal.SIGKILL)
                    except ProcessLookupError:
                        pass
                done_local = True
        else:
            subprocess.call(ssh_argv(h.address(), h.port, remote_command(user, prog)))
    print("Done killing")
    return 0


if __name__ == "__main__":
    sys.exit(main())
