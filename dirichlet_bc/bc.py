from heatflow_b200.dirichlet_bc.bc import RowDirichletBC  # noqa: F401
