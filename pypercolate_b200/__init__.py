"""pypercolate_b200 -- B200-native Newman-Ziff bond percolation (drop-in for
the hot path of andsor/pypercolate)."""
