"""sixdgs-b200: B200-native implementation of the 6DGS single-query pose-estimation hot path.
(placeholder -- filled in below)"""
