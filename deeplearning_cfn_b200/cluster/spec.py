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


PSEUDO_PARAMETERS = ("AWS::Region", "AWS::StackName", "AWS::StackId", "AWS::AccountId", "AWS::NoValue", "AWS::NotificationARNs",
                     "DLCFN::Region", "DLCFN::StackName", "DLCFN::StackId")
_TOP_LEVEL = ("TemplateFormatVersion", "AWSTemplateFormatVersion", "Description", "Metadata", "Parameters", "Mappings", "Conditions",
              "Resources", "Outputs")
_INTRINSICS = ("Ref", "Fn::FindInMap", "Fn::Join", "Fn::If", "Fn::Equals", "Fn::GetAtt")


def validate_template(template) -> dict:
    """Static check of a template, the counterpart of `aws cloudformation validate-template` (the stack scripts of the reference
    rely on CloudFormation rejecting a broken template before anything is created): known top-level sections, a Resources
    section whose entries have a Type, parameter declarations with consistent defaults, and every `Ref` / `Fn::GetAtt` /
    `Fn::FindInMap` / `DependsOn` / condition name pointing at something the template declares.  Returns the summary that call
    prints ({"Parameters": [{ParameterKey, DefaultValue, NoEcho, Description}], "Description", "Capabilities"}); raises
    TemplateError listing EVERY problem found."""
    t = load_template(template) if isinstance(template, str) else template
    errs: list[str] = []
    if not isinstance(t, dict):
        raise TemplateError("a template is a JSON object")
    for k in t:
        if k not in _TOP_LEVEL:
            errs.append(f"unknown top-level section {k!r}")
    params, maps, conds = t.get("Parameters", {}), t.get("Mappings", {}), t.get("Conditions", {})
    res, outs = t.get("Resources"), t.get("Outputs", {})
    if not isinstance(res, dict) or not res:
        errs.append("the Resources section is required and must declare at least one resource")
        res = {}
    for name, d in params.items():
        if not isinstance(d, dict) or "Type" not in d:
            errs.append(f"parameter {name}: Type is required")
            continue
        if d["Type"] not in ("String", "Number", "CommaDelimitedList") and not str(d["Type"]).startswith(("AWS::", "List<", "DLCFN::")):
            errs.append(f"parameter {name}: unknown Type {d['Type']!r}")
        if "Default" in d:
            try:
                resolve_parameters({"Parameters": {name: d}}, {})
            except TemplateError as e:
                errs.append(f"parameter {name}: its Default violates its own constraints ({e})")
    for name, r in res.items():
        if not isinstance(r, dict) or not isinstance(r.get("Type"), str):
            errs.append(f"resource {name}: Type is required")
            continue
        dep = r.get("DependsOn", [])
        for d in ([dep] if isinstance(dep, str) else dep):
            if d not in res:
                errs.append(f"resource {name}: DependsOn names {d!r}, which is not a resource")
        if "Condition" in r and r["Condition"] not in conds:
            errs.append(f"resource {name}: unknown condition {r['Condition']!r}")

    def walk(node, where: str) -> None:
        if isinstance(node, list):
            for i, x in enumerate(node):
                walk(x, f"{where}[{i}]")
            return
        if not isinstance(node, dict):
            return
        if len(node) == 1:
            (k, v), = node.items()
            if k == "Ref":
                if not (isinstance(v, str) and (v in params or v in res or v in PSEUDO_PARAMETERS)):
                    errs.append(f"{where}: Ref to {v!r}, which is neither a parameter, a resource nor a pseudo parameter")
                return
            if k == "Fn::GetAtt":
                target = v[0] if isinstance(v, list) and v else (v.split(".", 1)[0] if isinstance(v, str) else None)
                if target not in res:
                    errs.append(f"{where}: Fn::GetAtt on {target!r}, which is not a resource")
                return
            if k == "Fn::FindInMap":
                if not (isinstance(v, list) and len(v) == 3):
                    errs.append(f"{where}: Fn::FindInMap takes [map, key, key]")
                else:
                    if isinstance(v[0], str) and v[0] not in maps:
                        errs.append(f"{where}: Fn::FindInMap on {v[0]!r}, which is not a mapping")
                    elif isinstance(v[0], str) and isinstance(v[1], str) and v[1] not in maps[v[0]]:
                        errs.append(f"{where}: mapping {v[0]} has no key {v[1]!r}")
                    elif isinstance(v[0], str) and isinstance(v[1], str) and isinstance(v[2], str) and v[2] not in maps[v[0]][v[1]]:
                        errs.append(f"{where}: mapping {v[0]}/{v[1]} has no key {v[2]!r}")
                    walk(v, where + "/Fn::FindInMap")
                return
            if k == "Fn::If":
                if isinstance(v, list) and v and isinstance(v[0], str) and v[0] not in conds:
                    errs.append(f"{where}: Fn::If on unknown condition {v[0]!r}")
                walk(v[1:] if isinstance(v, list) else v, where + "/Fn::If")
                return
            if k.startswith("Fn::") and k not in _INTRINSICS:
                errs.append(f"{where}: {k} is not supported by this engine (supported: {', '.join(_INTRINSICS)})")
        for k, v in node.items():
            walk(v, f"{where}/{k}")

    for name, c in conds.items():
        walk(c, f"Conditions/{name}")
    for name, r in res.items():
        if isinstance(r, dict):
            walk(r.get("Properties", {}), f"Resources/{name}/Properties")
            walk(r.get("Metadata", {}), f"Resources/{name}/Metadata")
    for name, o in outs.items():
        if not isinstance(o, dict) or "Value" not in o:
            errs.append(f"output {name}: Value is required")
        else:
            walk(o["Value"], f"Outputs/{name}/Value")
    if errs:
        raise TemplateError("template validation failed:\n  " + "\n  ".join(errs))
    return {"Parameters": [{"ParameterKey": k, "DefaultValue": str(d["Default"]) if "Default" in d else None,
                            "NoEcho": str(d.get("NoEcho", "false")).lower() == "true", "Description": d.get("Description", "")}
                           for k, d in params.items()],
            "Description": t.get("Description", ""),
            "Capabilities": ["CAPABILITY_NAMED_IAM"] if any("Role" in str(r.get("Type", "")) or "Function" in str(r.get("Type", ""))
                                                           for r in res.values() if isinstance(r, dict)) else []}


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
        validate_template(t)              # what CloudFormation does before it creates anything
        p = resolve_parameters(t, given)
        st = t.get("Mappings", {}).get("StackTimeouts", {}).get("Timeouts", {})
        known = {"WorkerCount", "InstanceType", "ImageType", "EFSMountPoint", "EFSFileSystemId"}
        return cls(stack_name=stack_name, template=t, parameters=p, worker_count=int(p.get("WorkerCount", 1)),
                   instance_type=p.get("InstanceType", "b200.1gpu"), image_type=p.get("ImageType", "B200Torch"),
                   shared_mount=p.get("EFSMountPoint", "myEFSvolume"), shared_store_id=p.get("EFSFileSystemId", ""),
                   wait_timeout=float(st.get("WaitCondition", 3300)), master_launch_timeout=float(st.get("MasterLaunch", 600)),
                   extras={k: v for k, v in p.items() if k not in known})
