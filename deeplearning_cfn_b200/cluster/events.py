"""Scaling-group notifications and the function that reacts to them.

Reference: `cfn-lambda_function/lambda_function.py` - an SNS-triggered Lambda that (a) tells the master
agent when each Auto Scaling group is complete (`asg-setup` message, :51-62,117-119), (b) signals
CloudFormation for the master group so that worker creation may start (:121-128), (c) freezes
membership by suspending ReplaceUnhealthy (:129-132) and (d) on launch errors SHRINKS the worker
group's desired capacity to what actually launched, provided the minimum is met (:142-169) - the
provision-time elasticity described in README.md:49.

Intended behaviour is reproduced, the latent bugs are not: dispatch is a dict lookup (the reference
`eval`s a function name, :31) and the terminate-error handler accepts its message (the reference
forgot the parameter, :187).
"""
from __future__ import annotations

import json
import logging
import os
import time
import uuid
from collections import namedtuple

from .controlplane import ControlPlane

LOG = logging.getLogger("dlcfn-function")
ASGInstanceCount = namedtuple("ASGInstanceCount", ["min", "desired", "max", "launched"])

LAUNCH = "autoscaling:EC2_INSTANCE_LAUNCH"
LAUNCH_ERROR = "autoscaling:EC2_INSTANCE_LAUNCH_ERROR"
TERMINATE = "autoscaling:EC2_INSTANCE_TERMINATE"
TERMINATE_ERROR = "autoscaling:EC2_INSTANCE_TERMINATE_ERROR"
TEST = "autoscaling:TEST_NOTIFICATION"


def make_notification(event: str, group: str, instance_id: str, zone: str = "nvswitch-0", status_code: str = "",
                      status_message: str = "") -> dict:
    """Same envelope as an SNS-delivered Auto Scaling notification (lambda_function.py:24-35,97-101)."""
    msg = {"Event": event, "AutoScalingGroupName": group, "EC2InstanceId": instance_id,
           "Details": {"Availability Zone": zone}, "StartTime": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),
           "RequestId": uuid.uuid4().hex, "StatusCode": status_code, "StatusMessage": status_message}
    return {"Records": [{"Sns": {"Message": json.dumps(msg)}}]}


def _cp() -> ControlPlane:
    return ControlPlane(os.environ["AWS_DL_STACK_ID"])


def get_instance_count(cp: ControlPlane, group_name: str) -> ASGInstanceCount:
    """Healthy instances that are InService or Pending (lambda_function.py:67-92)."""
    g = cp.describe_group(group_name)
    healthy = sum(1 for i in g["Instances"]
                  if i["HealthStatus"] == "Healthy" and i["LifecycleState"] in ("InService", "Pending"))
    return ASGInstanceCount(min=g["MinSize"], max=g["MaxSize"], desired=g["DesiredCapacity"], launched=healthy)


def send_asg_success(cp: ControlPlane, status: str, asg: str, counts: ASGInstanceCount) -> None:
    body = dict(counts._asdict(), status=status.lower(), asg=asg, event="asg-setup")
    LOG.info("sending message to master queue: %s", json.dumps(body))
    cp.queue(os.environ["AWS_DL_MASTER_SQS_URL"]).send(body)


def _classify(group_name: str) -> str | None:
    if group_name and "WorkerAutoScalingGroup" in group_name:
        return "WorkerAutoScalingGroup"
    if group_name and "MasterAutoScalingGroup" in group_name:
        return "MasterAutoScalingGroup"
    return None


def on_instance_launch(message: dict) -> None:
    cp = _cp()
    name = message["AutoScalingGroupName"]
    kind = _classify(name)
    if kind is None:
        LOG.warning("unknown scaling group in %s", message)
        return
    counts = get_instance_count(cp, name)
    if counts.launched == counts.desired:
        send_asg_success(cp, "SUCCESS", name, counts)
        if kind == "MasterAutoScalingGroup":
            cp.signal_resource(kind, message["EC2InstanceId"], "SUCCESS")
        else:
            cp.suspend_processes(name, ["ReplaceUnhealthy"])


def on_instance_launch_error(message: dict) -> None:
    cp = _cp()
    name = message["AutoScalingGroupName"]
    counts = get_instance_count(cp, name)
    LOG.warning("launch error in %s: %s %s", name, message.get("StatusCode"), message.get("StatusMessage"))
    # shrink only if at least the requested minimum is there
    if counts.launched >= counts.min:
        cp.set_desired_capacity(name, counts.launched)
        cp.suspend_processes(name, ["ReplaceUnhealthy"])
        send_asg_success(cp, "SUCCESS", name, counts._replace(desired=counts.launched))


def on_instance_terminate(message: dict) -> None:
    LOG.info("instance %s of %s terminated", message.get("EC2InstanceId"), message.get("AutoScalingGroupName"))


def on_instance_terminate_error(message: dict) -> None:
    LOG.warning("terminate error for %s of %s", message.get("EC2InstanceId"), message.get("AutoScalingGroupName"))


def do_nothing(message: dict) -> None:
    LOG.info("ignored notification: %s", message.get("Event"))


_HANDLERS = {LAUNCH: on_instance_launch, LAUNCH_ERROR: on_instance_launch_error, TERMINATE: on_instance_terminate,
             TERMINATE_ERROR: on_instance_terminate_error, TEST: do_nothing}


def lambda_handler(event: dict, context=None):
    """Entry point with the reference's signature.  Needs AWS_DL_STACK_ID (stack directory) and
    AWS_DL_MASTER_SQS_URL (master queue name) in the environment (lambda_function.py:52,125)."""
    message = json.loads(event["Records"][0]["Sns"]["Message"])
    handler = _HANDLERS.get(message.get("Event"), do_nothing)
    handler(message)
    return message
