"""Stack templates and their parameters - the infrastructure-as-code tier.

Reference: `cfn-template/deeplearning.template` (Parameters :4-108, Mappings :113-178, Resources
:180-901, Outputs :903-926) and the Mask-RCNN variants (`mask-rcnn-cfn.yaml`, `private-mask-rcnn-cfn.yaml`).
The same declarative shape is kept - Parameters with Type/Default/AllowedValues/AllowedPattern/Min,
Mappings, Resources, Outputs, `Ref` / `Fn::FindInMap` / `Fn::Join` intrinsics - but the resources
describe ONE NVSwitch box: scaling groups of GPU *slots* instead of EC2 instances, spool-directory
queues instead of SQS, an in-process function instead of Lambda, a shared store (peer-HBM staging
arena + a directory) instead of EFS.
"""
from __future__ import annotations

import json
import os
import re
from dataclasses import dataclass, field

TEMPLATE_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "templates")

IMAGE_TYPES = ("B200Torch", "B200TorchConda", "CPUGloo")
INSTANCE_TYPES = ("b200.1gpu", "b200.2gpu", "b200.4gpu", "b200.8gpu", "cpu.large")
GPUS_PER_INSTANCE = {"b200.1gpu": 1, "b200.2gpu": 2, "b200.4gpu": 4, "b200.8gpu": 8, "cpu.large": 0}


class TemplateError(ValueError):
    pass


def load_template(name_or_path: str) -> dict:
    path = name_or_path
    if not os.path.exists(path):
        path = os.path.join(TEMPLATE_DIR, name_or_path)
        if not os.path.exists(path):
            for ext in (".template", ".json"):
                if os.path.exists(path + ext):
                    path = path + ext
                    break
    if not os.path.exists(path):
        raise TemplateError(f"template not found: {name_or_path}")
    with open(path) as f:
        return json.load(f)


def resolve_parameters(template: dict, given: dict) -> dict:
    """CloudFormation-style parameter validation (AllowedValues / AllowedPattern / Min/Max / Default)."""
    out = {}
    decl = template.get("Parameters", {})
    unknown = set(given) - set(decl)
    if unknown:
        raise TemplateError(f"unknown parameters: {sorted(unknown)}")
    for name, d in decl.items():
        if name in given:
            v = given[name]
        elif "Default" in d:
            v = d["Default"]
        else:
            raise TemplateError(f"parameter {name} is required")
        typ = d.get("Type", "String")
        if typ == "Number":
            try:
                v = int(v)
            except (TypeError, ValueError):
                raise TemplateError(f"parameter {name} must be a number, got {v!r}")
            if "MinValue" in d and v < int(d["MinValue"]):
                raise TemplateError(f"parameter {name}={v} below MinValue {d['MinValue']}")
            if "MaxValue" in d and v > int(d["MaxValue"]):
                raise TemplateError(f"parameter {name}={v} above MaxValue {d['MaxValue']}")
        else:
            v = str(v)
            if "AllowedValues" in d and v not in d["AllowedValues"]:
                raise TemplateError(f"parameter {name}={v!r} not in {d['AllowedValues']}: "
                                    f"{d.get('ConstraintDescription', '')}")
            if "AllowedPattern" in d and not re.fullmatch(d["AllowedPattern"], v):
                raise TemplateError(f"parameter {name}={v!r} does not match {d['AllowedPattern']}: "
                                    f"{d.get('ConstraintDescription', '')}")
            if "MinLength" in d and len(v) < int(d["MinLength"]):
                raise TemplateError(f"parameter {name} shorter than MinLength {d['MinLength']}")
            if "MaxLength" in d and len(v) > int(d["MaxLength"]):
                raise TemplateError(f"parameter {name} longer than MaxLength {d['MaxLength']}")
        out[name] = v
    return out


def evaluate(node, params: dict, mappings: dict, pseudo: dict, attrs: dict | None = None):
    """Evaluate Ref / Fn::FindInMap / Fn::Join / Fn::If / Fn::Equals / Fn::GetAtt inside a template value."""
    attrs = attrs or {}
    if isinstance(node, dict):
        if len(node) == 1:
            (k, v), = node.items()
            if k == "Ref":
                if v in params:
                    return params[v]
                if v in pseudo:
                    return pseudo[v]
                if v in attrs:
                    return attrs[v].get("Ref", v)
                raise TemplateError(f"unresolved Ref {v}")
            if k == "Fn::FindInMap":
                m, k1, k2 = [evaluate(x, params, mappings, pseudo, attrs) for x in v]
                try:
                    return mappings[m][k1][k2]
                except KeyError:
                    raise TemplateError(f"FindInMap miss: {m}/{k1}/{k2}")
            if k == "Fn::Join":
                sep, parts = v
                return sep.join(str(evaluate(p, params, mappings, pseudo, attrs)) for p in parts)
            if k == "Fn::Equals":
                a, b = [evaluate(x, params, mappings, pseudo, attrs) for x in v]
                return str(a) == str(b)
            if k == "Fn::If":
                c, a, b = v
                cond = evaluate(c, params, mappings, pseudo, attrs) if not isinstance(c, str) else pseudo["__conditions__"][c]
                return evaluate(a if cond else b, params, mappings, pseudo, attrs)
            if k == "Fn::GetAtt":
                res, att = v
                try:
                    return attrs[res][att]
                except KeyError:
                    raise TemplateError(f"unresolved GetAtt {res}.{att}")
        return {k: evaluate(v, params, mappings, pseudo, attrs) for k, v in node.items()}
    if isinstance(node, list):
        return [evaluate(x, params, mappings, pseudo, attrs) for x in node]
    return node


@dataclass
class ClusterSpec:
    """Resolved view of a stack request (what `launch_cluster` consumes)."""
    stack_name: str
    template: dict
    parameters: dict
    worker_count: int = 1            # WorkerCount: workers IN ADDITION to the master (deeplearning.template:11-17)
    instance_type: str = "b200.1gpu"
    image_type: str = "B200Torch"
    shared_mount: str = "myEFSvolume"   # EFSMountPoint (deeplearning.template:100-107)
    shared_store_id: str = ""           # EFSFileSystemId: reuse an existing store when set
    wait_timeout: float = 3300.0        # WaitCondition (deeplearning.template:174)
    master_launch_timeout: float = 600.0
    extras: dict = field(default_factory=dict)

    @property
    def gpus_per_instance(self) -> int:
        return GPUS_PER_INSTANCE[self.instance_type]

    @classmethod
    def from_template(cls, stack_name: str, template="deeplearning.template", **given) -> "ClusterSpec":
        if not re.fullmatch(r"[A-Za-z][A-Za-z0-9-]{0,127}", stack_name):
            raise TemplateError("stack name must match [A-Za-z][A-Za-z0-9-]*")
        t = load_template(template) if isinstance(template, str) else template
        p = resolve_parameters(t, given)
        st = t.get("Mappings", {}).get("StackTimeouts", {}).get("Timeouts", {})
        known = {"WorkerCount", "InstanceType", "ImageType", "EFSMountPoint", "EFSFileSystemId"}
        return cls(stack_name=stack_name, template=t, parameters=p, worker_count=int(p.get("WorkerCount", 1)),
                   instance_type=p.get("InstanceType", "b200.1gpu"), image_type=p.get("ImageType", "B200Torch"),
                   shared_mount=p.get("EFSMountPoint", "myEFSvolume"), shared_store_id=p.get("EFSFileSystemId", ""),
                   wait_timeout=float(st.get("WaitCondition", 3300)), master_launch_timeout=float(st.get("MasterLaunch", 600)),
                   extras={k: v for k, v in p.items() if k not in known})
