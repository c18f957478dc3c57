"""suffix_b200 -- B200-native suffix array / LCP construction behind the
BurntSushi/suffix `SuffixTable` API."""
