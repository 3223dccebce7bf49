from typing import Any
Adj = Any
Size = Any
