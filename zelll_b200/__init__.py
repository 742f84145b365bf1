"""zelll-b200: a B200-native cell-list engine behind zelll's API (hot path only)."""
