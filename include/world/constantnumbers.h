/* world-b200 drop-in for externs/WORLD_v2/src/world/constantnumbers.h:13-41: the numeric
 * constants of the WORLD API, same names and values, for callers that include the header. */
#ifndef WORLD_CONSTANT_NUMBERS_H_
#define WORLD_CONSTANT_NUMBERS_H_
namespace world {
const double kPi = 3.1415926535897932384;
const double kLog2 = 0.69314718055994529;
const double kEps = 0.00000000000000022204460492503131;
const double kMySafeGuardMinimum = 0.000000000001;
const double kMaximumValue = 100000.0;
/* F0 */
const double kFloorF0 = 71.0;
const double kCeilF0 = 800.0;
const double kDefaultF0 = 500.0;
const double kFloorF0StoneMask = 40.0;
const double kFloorF0D4C = 47.0;
/* windows */
const int kHanning = 1;
const int kBlackman = 2;
/* aperiodicity */
const double kFrequencyInterval = 3000.0;
const double kUpperLimit = 15000.0;
const double kThreshold = 0.85;
/* codec (mel axis) */
const double kM0 = 1127.01048;
const double kF0 = 700.0;
const double kFloorFrequency = 40.0;
const double kCeilFrequency = 20000.0;
}  // namespace world
#endif
