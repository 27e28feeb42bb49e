from .bc import RowDirichletBC, resolve_last_wins
