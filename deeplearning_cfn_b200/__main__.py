"""`python -m deeplearning_cfn_b200 <command>` - the user-facing CLI.

Control plane (counterparts of `aws cloudformation ...` as used by mask-rcnn-stack.sh:62-92):
    validate-template, create-stack / launch-cluster, list-stacks, describe-stacks, describe-stack-events, delete-stack,
    prepare-bucket, node-exec
Job plane:
    launch (tools/launch.py compatible), kill, generate-trainer, run-example, bandwidth, benchmark, im2rec, rec2idx, parse-log, diagnose
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys


def _params(items: list[str]) -> dict:
    out = {}
    for it in items or []:
        if it.startswith("ParameterKey="):
            k, v = it.split(",", 1)
            out[k.split("=", 1)[1]] = v.split("=", 1)[1] if "=" in v else ""
        elif "=" in it:
            k, v = it.split("=", 1)
            out[k] = v
    return out


def cmd_create_stack(a) -> int:
    from .cluster.stack import launch_cluster
    template = a.template_body[7:] if a.template_body.startswith("file://") else a.template_body
    params = {k: v for k, v in _params(a.parameters).items()}
    if a.no_wait:
        # detach: the caller polls list-stacks exactly like mask-rcnn-stack.sh does
        argv = [sys.executable, "-m", "deeplearning_cfn_b200", "create-stack", "--stack-name", a.stack_name,
                "--template-body", a.template_body, "--backend", a.backend]
        if a.home:
            argv += ["--home", a.home]
        if a.capacity_gpus is not None:
            argv += ["--capacity-gpus", str(a.capacity_gpus)]
        argv += ["--parameters"] + [f"{k}={v}" for k, v in params.items()]
        subprocess.Popen(argv, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, start_new_session=True)
        print(json.dumps({"StackId": a.stack_name}))
        return 0
    h = launch_cluster(a.stack_name, template, home=a.home, backend=a.backend, capacity_gpus=a.capacity_gpus,
                       poll_interval=a.poll_interval, **params)
    print(json.dumps({"StackId": h.stack_dir, "StackStatus": h.status, "Outputs": h.outputs, "Hostfile": h.hostfile,
                      "Env": h.export_env(), "Workers": h.workers}, indent=1))
    return 0 if h.status == "CREATE_COMPLETE" else 1


def cmd_validate_template(a) -> int:
    from .cluster.spec import validate_template
    template = a.template_body[7:] if a.template_body.startswith("file://") else a.template_body
    print(json.dumps(validate_template(template), indent=1))
    return 0


def cmd_list_stacks(a) -> int:
    from .cluster.stack import list_stacks
    for s in list_stacks(a.home):
        if a.stack_status_filter and s["StackStatus"] not in a.stack_status_filter:
            continue
        print(json.dumps(s))
    return 0


def cmd_describe_stacks(a) -> int:
    from .cluster.stack import describe_stack
    h = describe_stack(a.stack_name, a.home)
    print(json.dumps({"StackName": h.stack_name, "StackStatus": h.status, "Outputs": h.outputs, "Hostfile": h.hostfile,
                      "HostsFile": h.hosts_file, "Env": h.export_env(), "Workers": h.workers}, indent=1))
    return 0


def cmd_describe_stack_events(a) -> int:
    """Newest first, like `aws cloudformation describe-stack-events` - where a user of the reference looks when a stack rolls
    back (which resource failed and why: a master that could not launch, an agent that timed out, a shrink)."""
    import datetime
    from .cluster.controlplane import ControlPlane
    from .cluster.stack import DEFAULT_HOME, StackError
    sdir = os.path.join(a.home or DEFAULT_HOME, "stacks", a.stack_name)
    if not os.path.exists(os.path.join(sdir, "stack.json")):
        raise StackError(f"no such stack: {a.stack_name}")
    evs = ControlPlane(sdir).stack().get("events", [])
    out = [{"Timestamp": datetime.datetime.fromtimestamp(e["ts"], datetime.timezone.utc).isoformat(timespec="milliseconds"),
            "StackName": a.stack_name, "LogicalResourceId": e["id"], "ResourceStatus": e["status"],
            "ResourceStatusReason": e.get("reason", "")} for e in reversed(evs)]
    print(json.dumps({"StackEvents": out}, indent=1))
    return 0


def cmd_delete_stack(a) -> int:
    from .cluster.stack import delete_stack
    delete_stack(a.stack_name, a.home)
    return 0


def cmd_prepare_bucket(a) -> int:
    from .cluster.staging import prepare_bucket
    print(json.dumps(prepare_bucket(a.bucket, a.prefix, a.num_images, a.image_size), indent=1))
    return 0


def cmd_node_exec(a) -> int:
    from .launch.transport import node_exec
    return node_exec(a.host, a.command)


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    # job-plane commands own their argument parsing
    passthrough = {"launch": "launch.launcher", "kill": "launch.kill", "generate-trainer": "launch.tf_ps_scripts",
                   "run-example": "examples.run_example", "bandwidth": "examples.bandwidth_measure",
                   "benchmark": "examples.benchmark", "mpirun": "launch.mpirun", "im2rec": "utils.im2rec",
                   "parse-log": "utils.parse_log", "rec2idx": "utils.recordio", "diagnose": "utils.diagnose"}
    if argv and argv[0] in passthrough:
        import importlib
        mod = importlib.import_module("." + passthrough[argv[0]], __package__)
        try:
            return int(mod.main(argv[1:]) or 0)
        except (FileNotFoundError, IsADirectoryError, NotADirectoryError, PermissionError) as e:
            # a missing / unreadable hostfile, log or list file is a usage error, not a crash
            print(f"{argv[0]}: error: {e.strerror}: {e.filename}", file=sys.stderr)
            return 1
        except Exception as e:
            from .launch.hostfile import HostfileError
            if isinstance(e, HostfileError):      # so is a hostfile line that cannot be read
                print(f"{argv[0]}: error: {e}", file=sys.stderr)
                return 1
            raise

    ap = argparse.ArgumentParser(prog="dlcfn", formatter_class=argparse.RawDescriptionHelpFormatter,
                                 epilog="job-plane commands (each with its own --help): " + ", ".join(sorted(passthrough)))
    sub = ap.add_subparsers(dest="cmd", required=True)
    p = sub.add_parser("validate-template", help="static check of a template (aws cloudformation validate-template)")
    p.add_argument("--template-body", default="deeplearning.template")
    p.set_defaults(fn=cmd_validate_template)
    for name in ("create-stack", "launch-cluster"):
        p = sub.add_parser(name)
        p.add_argument("--stack-name", default="dlcfn")
        p.add_argument("--template-body", default="deeplearning.template")
        p.add_argument("--region", default="b200-box-1")
        p.add_argument("--capabilities", default="")
        p.add_argument("--parameters", nargs="*", default=[])
        p.add_argument("--home", default=None)
        p.add_argument("--backend", default="process", choices=["process", "thread"])
        p.add_argument("--capacity-gpus", type=int, default=None)
        p.add_argument("--poll-interval", type=float, default=0.05)
        p.add_argument("--no-wait", action="store_true")
        p.set_defaults(fn=cmd_create_stack)
    p = sub.add_parser("list-stacks")
    p.add_argument("--stack-status-filter", nargs="*", default=[])
    p.add_argument("--home", default=None)
    p.add_argument("--region", default=None)
    p.set_defaults(fn=cmd_list_stacks)
    for name, fn in (("describe-stacks", cmd_describe_stacks), ("describe-stack-events", cmd_describe_stack_events),
                     ("delete-stack", cmd_delete_stack)):
        p = sub.add_parser(name)
        p.add_argument("--stack-name", required=True)
        p.add_argument("--home", default=None)
        p.add_argument("--region", default=None)
        p.set_defaults(fn=fn)
    p = sub.add_parser("prepare-bucket")
    p.add_argument("--bucket", required=True)
    p.add_argument("--prefix", default="mask-rcnn/b200")
    p.add_argument("--num-images", type=int, default=64)
    p.add_argument("--image-size", type=int, default=128)
    p.set_defaults(fn=cmd_prepare_bucket)
    p = sub.add_parser("package-event-function")
    p.add_argument("--output", default="dl_cfn_setup_lambda.zip")
    p.set_defaults(fn=lambda a: (print(__import__("deeplearning_cfn_b200.cluster.staging", fromlist=["x"])
                                       .package_event_function(a.output)), 0)[1])
    p = sub.add_parser("node-exec")
    p.add_argument("host")
    p.add_argument("command", nargs=argparse.REMAINDER)
    p.set_defaults(fn=cmd_node_exec)
    a = ap.parse_args(argv)
    try:
        return a.fn(a)
    except Exception as e:
        print(f"error: {e}", file=sys.stderr)
        return 1


if __name__ == "__main__":
    sys.exit(main())
