"""Settings: attribute-style view of input/paramfile.yml plus CLI overrides.

Same behaviour as the reference's dgfem/settings.py:3-72 (keys with spaces become
underscored attributes, nested dicts become nested Settings, update_setting creates missing
leaves, update_settings maps the CLI keywords onto setting paths) so that paramfiles and
command lines written for `python -m dgfem` work unchanged.

B200-specific, optional keys (all have defaults, unknown to the reference and ignored by it):
  solver.b200.gs_mode          lexicographic (default; bit-for-bit the reference's sweep order)
                               | redblack (2-colour multicolour sweep)
  solver.b200.check_residual   True (default; keeps the smoother's 1e-6 early exit / 1e10 abort)
"""
import os


def load_params(path=None):
    """input/__init__.py:4-7: the paramfile is looked up relative to the working directory."""
    from yaml import safe_load
    path = path or os.path.join(os.getcwd(), "input", "paramfile.yml")
    with open(path, "r") as f:
        return safe_load(f)


def update_parameters(params, keys, value):
    """input/__init__.py:9-12."""
    node = params
    for k in keys[:-1]:
        node = node[k]
    node[keys[-1]] = value


class Settings:
    def __init__(self, params):
        for key, value in params.items():
            name = key.replace(" ", "_")
            setattr(self, name, Settings(value) if isinstance(value, dict) else value)

    def get(self, path, default=None):
        node = self
        for k in path.split("."):
            if not hasattr(node, k):
                return default
            node = getattr(node, k)
        return node

    def _attribute_exists(self, attribute_path):
        return self.get(attribute_path, _MISSING) is not _MISSING

    def update_setting(self, attribute_path, new_value):
        keys = attribute_path.split(".")
        node = self
        for k in keys[:-1]:
            if not hasattr(node, k):
                setattr(node, k, Settings({}))
            node = getattr(node, k)
        setattr(node, keys[-1], new_value)

    # CLI keyword -> setting path (dgfem/settings.py:46-72)
    _CLI_MAP = (
        ("grid_folder", "grid.folder"), ("grid_file", "grid.filename"),
        ("p_grid", "grid.polynomial_degree"), ("p_solution", "solution.polynomial_degree"),
        ("manufactured_solution", "solution.manufactured_solution"),
        ("solution_polynomial_degree_u", "solution.u.polynomial_degree"),
        ("solution_polynomial_degree_p", "solution.p.polynomial_degree"),
        ("solution_ordering", "solution.ordering"),
        ("problem_governing_equations", "problem.governing_equation(s)"),
        ("problem_kinematic_viscosity", "problem.kinematic_viscosity"),
        ("SIP_penalty_parameter", "problem.SIP_penalty_parameter"),
        ("SIP_penalty_parameter_multiplier", "problem.SIP_penalty_parameter multiplier"),
        ("velocity_penalty_parameter", "problem.velocity_penalty_parameter"),
        ("exact_solution_u", "problem.exact_solution.u"), ("exact_solution_v", "problem.exact_solution.v"),
        ("exact_solution_p", "problem.exact_solution.p"), ("exact_solution_tag", "problem.exact_solution.tag"),
        ("smoother", "solver.smoother"),
    )

    def update_settings(self, kwargs):
        for key, path in self._CLI_MAP:
            if kwargs.get(key):
                self.update_setting(path, kwargs.get(key))
        if kwargs.get("check_eigenvalues"):
            self.update_setting("solver.check_eigenvalues", True)
        if kwargs.get("check_condition_number"):
            self.update_setting("problem.check_condition_number", True)
        if kwargs.get("plot_sparsity_pattern"):
            self.update_setting("visualization.plot_sparsity_pattern", True)
        self.update_setting("solver.discretization", kwargs.get("discretization") or "dg")
        if kwargs.get("solve_finite_volume_method"):
            self.update_setting("solver.discretization", "fvm")

    def _validate_settings(self, settings):
        """dgfem/settings.py:23-36."""
        if settings.problem.type == "Poisson":
            assert settings.solution.ordering == "local"
        if settings.problem.type == "Stokes" and settings.solver.method == "multigrid":
            assert settings.solution.ordering == "global"
            assert settings.problem.multiply_inverse_mass_matrix is True


_MISSING = object()
