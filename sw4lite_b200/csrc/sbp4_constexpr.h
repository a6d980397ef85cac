// GENERATED from sbp4_tables.h (scripts/gen_sbp_constexpr.py) -- do not edit.
// The same tables as compile-time functions: with constant arguments (fully unrolled loops) the
// compiler folds the values into the instruction stream and drops the zero entries altogether.
// Used by k_closure_fast only when the runtime tables equal these built-in ones (api.cu checks).
#ifndef SW4B200_SBP4_CONSTEXPR_H
#define SW4B200_SBP4_CONSTEXPR_H
namespace sw4b200 {
__host__ __device__ constexpr double acof_c( int idx )
{
   switch( idx )
   {
   case 0: return 104.0 / 289.0;
   case 1: return 12.0 / 17.0;
   case 2: return -96.0 / 731.0;
   case 3: return -36.0 / 833.0;
   case 6: return -516.0 / 289.0;
   case 7: return -59.0 / 68.0;
   case 8: return 118.0 / 731.0;
   case 9: return 177.0 / 3332.0;
   case 12: return 312.0 / 289.0;
   case 13: return 2.0 / 17.0;
   case 14: return -16.0 / 731.0;
   case 15: return -6.0 / 833.0;
   case 18: return -104.0 / 289.0;
   case 19: return 3.0 / 68.0;
   case 20: return -6.0 / 731.0;
   case 21: return -9.0 / 3332.0;
   case 48: return -2476335.0 / 2435692.0;
   case 49: return 544521.0 / 4226642.0;
   case 50: return 1024279.0 / 6160868.0;
   case 51: return 181507.0 / 3510262.0;
   case 54: return 544521.0 / 1217846.0;
   case 55: return -1633563.0 / 4226642.0;
   case 56: return 1633563.0 / 3080434.0;
   case 57: return -544521.0 / 3510262.0;
   case 60: return 1024279.0 / 2435692.0;
   case 61: return 1633563.0 / 4226642.0;
   case 62: return -5380447.0 / 6160868.0;
   case 63: return 544521.0 / 3510262.0;
   case 66: return 181507.0 / 1217846.0;
   case 67: return -544521.0 / 4226642.0;
   case 68: return 544521.0 / 3080434.0;
   case 69: return -181507.0 / 3510262.0;
   case 96: return -16189.0 / 84966.0;
   case 97: return 2509879.0 / 12679926.0;
   case 98: return -687797.0 / 3080434.0;
   case 99: return 241309.0 / 10530786.0;
   case 100: return 5.0 / 6192.0;
   case 102: return 2509879.0 / 3653538.0;
   case 103: return -21510077.0 / 25359852.0;
   case 104: return 2565299.0 / 3080434.0;
   case 105: return 987685.0 / 21061572.0;
   case 106: return 815.0 / 151704.0;
   case 108: return -687797.0 / 1217846.0;
   case 109: return 2565299.0 / 4226642.0;
   case 110: return -3569115.0 / 3080434.0;
   case 111: return 2193521.0 / 3510262.0;
   case 112: return -7381.0 / 50568.0;
   case 114: return 241309.0 / 3653538.0;
   case 115: return 987685.0 / 25359852.0;
   case 116: return 2193521.0 / 3080434.0;
   case 117: return -2647979.0 / 3008796.0;
   case 118: return 28709.0 / 151704.0;
   case 120: return 5.0 / 2193.0;
   case 121: return 1630.0 / 372939.0;
   case 122: return -14762.0 / 90601.0;
   case 123: return 57418.0 / 309729.0;
   case 124: return -349.0 / 7056.0;
   case 144: return -9.0 / 3332.0;
   case 146: return 177.0 / 8428.0;
   case 148: return -1.0 / 49.0;
   case 149: return 1.0 / 392.0;
   case 151: return -12655.0 / 372939.0;
   case 152: return 40072.0 / 271803.0;
   case 153: return -14762.0 / 103243.0;
   case 154: return 1186.0 / 18963.0;
   case 155: return -1.0 / 144.0;
   case 156: return 177.0 / 3332.0;
   case 157: return 40072.0 / 372939.0;
   case 158: return -331815.0 / 362404.0;
   case 159: return 8065.0 / 14749.0;
   case 160: return 32555.0 / 303408.0;
   case 161: return 3.0 / 784.0;
   case 163: return -14762.0 / 124313.0;
   case 164: return 8065.0 / 12943.0;
   case 165: return -80793.0 / 103243.0;
   case 166: return 51269.0 / 101136.0;
   case 167: return -283.0 / 2352.0;
   case 168: return -48.0 / 833.0;
   case 169: return 18976.0 / 372939.0;
   case 170: return 32555.0 / 271803.0;
   case 171: return 51269.0 / 103243.0;
   case 172: return -247951.0 / 303408.0;
   case 173: return 1135.0 / 7056.0;
   case 174: return 6.0 / 833.0;
   case 175: return -1.0 / 177.0;
   case 176: return 9.0 / 2107.0;
   case 177: return -283.0 / 2401.0;
   case 178: return 1135.0 / 7056.0;
   case 179: return -47.0 / 1176.0;
   case 206: return -283.0 / 6321.0;
   case 207: return 381.0 / 2401.0;
   case 208: return -283.0 / 2352.0;
   case 209: return -11.0 / 7056.0;
   case 212: return 381.0 / 2107.0;
   case 213: return -1927.0 / 2401.0;
   case 214: return 381.0 / 784.0;
   case 215: return 403.0 / 2352.0;
   case 218: return -283.0 / 2107.0;
   case 219: return 1143.0 / 2401.0;
   case 220: return -577.0 / 784.0;
   case 221: return 1165.0 / 2352.0;
   case 224: return -11.0 / 6321.0;
   case 225: return 403.0 / 2401.0;
   case 226: return 1165.0 / 2352.0;
   case 227: return -5869.0 / 7056.0;
   case 232: return -1.0 / 8.0;
   case 233: return 1.0 / 6.0;
   case 261: return -2.0 / 49.0;
   case 262: return 1.0 / 6.0;
   case 263: return -1.0 / 8.0;
   case 267: return 8.0 / 49.0;
   case 268: return -5.0 / 6.0;
   case 269: return 1.0 / 2.0;
   case 273: return -6.0 / 49.0;
   case 274: return 1.0 / 2.0;
   case 275: return -3.0 / 4.0;
   case 280: return 1.0 / 6.0;
   case 281: return 1.0 / 2.0;
   case 287: return -1.0 / 8.0;
   case 316: return -1.0 / 24.0;
   case 317: return 1.0 / 6.0;
   case 322: return 1.0 / 6.0;
   case 323: return -5.0 / 6.0;
   case 328: return -1.0 / 8.0;
   case 329: return 1.0 / 2.0;
   case 335: return 1.0 / 6.0;
   case 371: return -1.0 / 24.0;
   case 377: return 1.0 / 6.0;
   case 383: return -1.0 / 8.0;
   default: return 0.0;
   }
}
__host__ __device__ constexpr double bope_c( int idx )
{
   switch( idx )
   {
   case 0: return -24.0 / 17.0;
   case 1: return -1.0 / 2.0;
   case 2: return 4.0 / 43.0;
   case 3: return 3.0 / 98.0;
   case 6: return 59.0 / 34.0;
   case 8: return -59.0 / 86.0;
   case 12: return -4.0 / 17.0;
   case 13: return 1.0 / 2.0;
   case 15: return -59.0 / 98.0;
   case 16: return 1.0 / 12.0;
   case 18: return -3.0 / 34.0;
   case 20: return 59.0 / 86.0;
   case 22: return -2.0 / 3.0;
   case 23: return 1.0 / 12.0;
   case 26: return -4.0 / 43.0;
   case 27: return 32.0 / 49.0;
   case 29: return -2.0 / 3.0;
   case 33: return -4.0 / 49.0;
   case 34: return 2.0 / 3.0;
   case 40: return -1.0 / 12.0;
   case 41: return 2.0 / 3.0;
   case 47: return -1.0 / 12.0;
   default: return 0.0;
   }
}
__host__ __device__ constexpr double ghcof_c( int idx )
{
   switch( idx )
   {
   case 0: return 12.0 / 17.0;
   default: return 0.0;
   }
}
} // namespace sw4b200
#endif
