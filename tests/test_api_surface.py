"""Drop-in check of the API surface: every function the reference's `jpc/__init__.py` re-exports is either provided
by jpc_b200 under the same name with the same parameters (names, order, positional vs keyword-only, literal defaults),
or is on the explicit out-of-scope list (PDM and the linear-network analytical tools: SURVEY 2 / DESIGN 8).
The reference side is a committed fixture (tests/golden/api_signatures.json, made by tests/golden/make_api_signatures.py
with `ast`; nothing is read from /root/reference at test time)."""
import inspect
import json
import os

OUT_OF_SCOPE = {
    # PDM ("predictive dendritic model") energy family -- a different energy, not on the north_star path
    "pdm_energy_fn", "compute_pdm_activity_grad", "compute_pdm_param_grads", "update_pdm_activities", "update_pdm_params",
    # closed-form theory of deep linear networks (jpc/_core/_analytical.py): analysis tools, restated only inside the
    # oracle as known-answer tests
    "linear_equilib_energy", "compute_linear_equilib_rescaling", "compute_linear_activity_hessian",
    "compute_linear_activity_solution", "compute_linear_equilib_energy_grads", "update_linear_equilib_energy_params",
}
# parameters this implementation adds on top of the reference's (all keyword, all optional)
EXTRA_OK = {"batch_global", "device"}


def _fixture():
    return json.load(open(os.path.join(os.path.dirname(__file__), "golden", "api_signatures.json")))


def test_every_exported_name_is_provided_or_explicitly_out_of_scope():
    import jpc_b200 as J
    fx = _fixture()
    missing = {n for n in fx["exported"] if not hasattr(J, n)}
    assert missing == OUT_OF_SCOPE


def test_signatures_match_the_reference():
    import jpc_b200 as J
    fx = _fixture()
    problems = []
    for name, ref in sorted(fx["signatures"].items()):
        if name in OUT_OF_SCOPE:
            continue
        sig = inspect.signature(getattr(J, name))
        mine = [(p.name, p.kind, p.default) for p in sig.parameters.values() if p.name not in EXTRA_OK]
        want = ref["params"]
        if [m[0] for m in mine] != [w[0] for w in want]:
            problems.append(f"{name}: parameters {[m[0] for m in mine]} != reference {[w[0] for w in want]} ({ref['file']})")
            continue
        for (pname, kind, default), (_, wkind, wdefault) in zip(mine, want):
            # the reference passes everything by keyword in its own code and tests; a keyword-only parameter must not have
            # become positional-only, and positional ones must still be positional
            if wkind == "positional" and kind not in (inspect.Parameter.POSITIONAL_OR_KEYWORD,):
                problems.append(f"{name}.{pname}: reference takes it positionally")
            if wdefault == "<required>":
                if default is not inspect.Parameter.empty:
                    problems.append(f"{name}.{pname}: required in the reference, optional here")
            elif wdefault == "<expr>":
                # constructed defaults (Heun(), PIDController(...)): here None stands for "the reference's default object"
                if default is inspect.Parameter.empty:
                    problems.append(f"{name}.{pname}: optional in the reference, required here")
            else:
                if default is inspect.Parameter.empty or repr(default) != wdefault and not _same_number(default, wdefault):
                    problems.append(f"{name}.{pname}: default {default!r} != reference {wdefault}")
    assert not problems, "\n".join(problems)


def _same_number(value, ref_repr):
    try:
        return float(value) == float(ref_repr)
    except (TypeError, ValueError):
        return False
